"""`import DenseNet; DenseNet.DenseNet()` -> sm_100a implementation (RegressionNetwork/train.py:13,39; test.py:13,29)."""
from emlight_b200.densenet import DenseNet  # noqa: F401
