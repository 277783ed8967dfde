"""Drop-in module name of the reference's Boost.Python extension (reference cpp/src/module.cpp:26-50,
cpp/CMakeLists.txt:55): reference scripts `import level_set_fusion_optimization as cpp` unchanged and get the
B200 implementation."""
from lsf_b200 import *  # noqa: F401,F403
from lsf_b200 import ops, telemetry, slavcheva, tsdf  # noqa: F401
