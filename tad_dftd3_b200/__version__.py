"""Version of the package (reference `__version__.py`)."""
__all__ = ["__version__"]

__version__ = "0.1.0"
