from pytorch_bayesiancnn_b200.modules import BBBConv2d  # layers/BBB/BBBConv.py:14
