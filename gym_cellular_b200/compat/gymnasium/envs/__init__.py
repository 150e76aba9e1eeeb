from . import registration
from .registration import register, make, make_vec, registry, spec
