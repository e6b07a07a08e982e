from pytorch_bayesiancnn_b200.modules import FlattenLayer, ModuleWrapper  # layers/misc.py
