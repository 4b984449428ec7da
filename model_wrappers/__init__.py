"""Drop-in module path of the reference's model_wrappers package (model_wrappers/model_loader.py)."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
