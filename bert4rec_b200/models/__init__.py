from .bert4rec_model import BERT4RecModel, History, StepMetrics, SPECIAL_TOKEN_IDS  # noqa: F401
from . import model_utils  # noqa: F401
