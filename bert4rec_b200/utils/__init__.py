from .utils import load_json_config, get_project_root, get_default_model_save_path  # noqa: F401
