"""from BabelViscoFDTD.tools.RayleighAndBHTE import ForwardSimple, InitCuda, ...
(TranscranialModeling/BabelIntegrationBASE.py:19, BabelIntegrationSingle.py:23, H317.py:5).
The Rayleigh integral, device selection and the bio-heat solver (BHTE / BHTEMultiplePressureFields, used by
ThermalModeling/CalculateTemperatureEffects.py:14) all run on the B200 library."""
from babelbrain_b200.rayleigh import (ForwardSimple, InitCuda, InitOpenCL, InitMetal, InitMLX,  # noqa: F401
                                      SpeedofSoundWater, GenerateFocusTx)
from babelbrain_b200.thermal import BHTE, BHTEMultiplePressureFields  # noqa: F401
