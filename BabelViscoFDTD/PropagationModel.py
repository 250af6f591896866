"""from BabelViscoFDTD.PropagationModel import PropagationModel
(TranscranialModeling/BabelIntegrationBASE.py:18, BabelIntegrationREMOPD.py:24)."""
from babelbrain_b200.propagation import PropagationModel  # noqa: F401
