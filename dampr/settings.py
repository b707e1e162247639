import sys
import dampr_b200.settings as _s
sys.modules[__name__] = _s
