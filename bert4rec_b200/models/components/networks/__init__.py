from .bert4rec_encoder import Bert4RecEncoder  # noqa: F401
