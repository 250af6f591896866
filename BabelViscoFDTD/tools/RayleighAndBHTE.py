"""from BabelViscoFDTD.tools.RayleighAndBHTE import ForwardSimple, InitCuda, ...
(TranscranialModeling/BabelIntegrationBASE.py:19, BabelIntegrationSingle.py:23, H317.py:5)."""
from babelbrain_b200.rayleigh import (ForwardSimple, InitCuda, InitOpenCL, InitMetal, InitMLX,  # noqa: F401
                                      SpeedofSoundWater, GenerateFocusTx, BHTE, BHTEMultiplePressureFields)
