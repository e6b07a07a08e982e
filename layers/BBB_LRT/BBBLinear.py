from pytorch_bayesiancnn_b200.modules import BBBLRTLinear as BBBLinear  # layers/BBB_LRT/BBBLinear.py:16
