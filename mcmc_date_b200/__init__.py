"""Import alias: the package sources live in `mcmc-date_b200/` (a directory name Python cannot
import directly); this shim puts that directory on the package path."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mcmc-date_b200"))
