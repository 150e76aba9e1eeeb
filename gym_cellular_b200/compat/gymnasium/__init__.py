"""Structural stand-in for the `gymnasium` package (fallback only).

The image this repo is built and run in has no `gymnasium` wheel and no network.
`gym_cellular_b200` subclasses the real gymnasium classes whenever
`import gymnasium` succeeds; only when it does not, the directory holding this
package is appended to the END of `sys.path`, so the real package always wins.

It provides the names the gym-cellular surface touches -- `Env`, `spaces`,
`vector.VectorEnv`, `envs.registration.{register, make, make_vec}` -- with
gymnasium's signatures.  It is also what lets the unmodified reference be
imported for golden-vector generation (tests/golden/make_golden.py): the
reference uses gymnasium only structurally (base class, space constructors,
`register`), never for arithmetic.
"""
from . import spaces
from .core import Env, Wrapper
from . import envs
from .envs.registration import register, make, make_vec, registry, spec
from . import vector

__version__ = "0.0.0+gym_cellular_b200.compat"
IS_COMPAT_STANDIN = True

__all__ = ["Env", "Wrapper", "spaces", "envs", "vector", "register", "make", "make_vec",
           "registry", "spec"]
