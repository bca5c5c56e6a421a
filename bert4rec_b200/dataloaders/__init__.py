"""Host-side data path: samplers, Cloze masking, batch layout (reference: bert4rec/dataloaders/)."""
from . import samplers  # noqa: F401
from . import dataloader_utils  # noqa: F401
from .preprocessors import BERT4RecPreprocessor  # noqa: F401
