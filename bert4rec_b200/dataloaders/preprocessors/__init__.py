from .bert4rec_preprocessor import BERT4RecPreprocessor  # noqa: F401

preprocessors_map = {"bert4rec": BERT4RecPreprocessor}


def get(identifier="bert4rec", **kwargs):
    if identifier in preprocessors_map:
        return preprocessors_map[identifier]
    raise ValueError(f"{identifier} is not known!")
