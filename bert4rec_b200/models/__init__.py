from .bert4rec_model import BERT4RecModel, History, StepMetrics, SPECIAL_TOKEN_IDS  # noqa: F401
from . import model_utils  # noqa: F401
from .model_wrapper import ModelWrapper  # noqa: F401
from .bert4rec_wrapper import BERT4RecModelWrapper  # noqa: F401
