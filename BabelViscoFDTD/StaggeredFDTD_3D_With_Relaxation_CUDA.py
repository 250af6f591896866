"""ListDevices() as the GUI's device picker calls it (BabelBrain/SelFiles/SelFiles.py:245-262)."""
from babelbrain_b200 import _capi


def ListDevices():
    return _capi.device_names()
