from pytorch_bayesiancnn_b200.modules import BBBLinear  # layers/BBB/BBBLinear.py:14
