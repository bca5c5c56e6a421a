"""Small path / config helpers (reference: bert4rec/utils/utils.py:14-40).  Unlike the reference, importing this
module does not require the VIRTUAL_ENV environment variable."""
import json
import pathlib

DEFAULT_MODEL_SAVE_PATH = pathlib.Path("saved_models")


def get_project_root() -> pathlib.Path:
    return pathlib.Path(__file__).resolve().parent.parent.parent


def get_default_model_save_path() -> pathlib.Path:
    return DEFAULT_MODEL_SAVE_PATH


def load_json_config(save_path) -> dict:
    """Loads a JSON file into a dict.  A bare config name such as ``"ml-1m_64"`` (or a path whose file does not exist
    but whose stem names one of the built-in encoder configs) resolves to ``bert4rec_b200.config.encoder_config``."""
    from bert4rec_b200 import config as _cfg
    p = pathlib.Path(save_path)
    if p.is_file():
        with open(p, "r") as f:
            return json.load(f)
    if p.stem in _cfg.available_configs():
        return _cfg.encoder_config(p.stem)
    raise FileNotFoundError(f"No config file exists at given path: {save_path}")
