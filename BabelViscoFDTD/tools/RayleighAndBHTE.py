"""from BabelViscoFDTD.tools.RayleighAndBHTE import ForwardSimple, InitCuda, ...
(TranscranialModeling/BabelIntegrationBASE.py:19, BabelIntegrationSingle.py:23, H317.py:5).
The Rayleigh integral and device selection run on the B200 library; the bio-heat functions are not part
of this path and come from the genuine package when it is installed (BabelViscoFDTD.upstream())."""
from babelbrain_b200.rayleigh import (ForwardSimple, InitCuda, InitOpenCL, InitMetal, InitMLX,  # noqa: F401
                                      SpeedofSoundWater, GenerateFocusTx)
from babelbrain_b200 import rayleigh as _r
from BabelViscoFDTD import upstream_attr as _up

BHTE = _up('tools.RayleighAndBHTE', 'BHTE') or _r.BHTE
BHTEMultiplePressureFields = _up('tools.RayleighAndBHTE', 'BHTEMultiplePressureFields') or _r.BHTEMultiplePressureFields
