"""softspoken_b200: B200-native voice-detector batch path of AVianEco/Softspoken."""
from . import spec  # noqa: F401

__version__ = "0.1.0"
