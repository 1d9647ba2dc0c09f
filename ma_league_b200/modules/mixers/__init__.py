from .qmix import QMixer
from .vdn import VDNMixer
