from .ranker import Ranker  # noqa: F401
from .recommender import Recommender  # noqa: F401
from .inference import InferenceDataloader  # noqa: F401
