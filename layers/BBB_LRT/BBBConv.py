from pytorch_bayesiancnn_b200.modules import BBBLRTConv2d as BBBConv2d  # layers/BBB_LRT/BBBConv.py:16
