"""`interpn.raw`: the 16 raw bindings of /root/reference/src/python.rs:55-292, served by interpn_b200.raw."""

from interpn_b200.raw import *  # noqa: F401,F403
from interpn_b200.raw import __all__  # noqa: F401
