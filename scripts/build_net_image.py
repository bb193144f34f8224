"""compile the Kuramoto network image (bench / scaling scripts) into the in-tree cache without a GPU; variants via
JITCDDE_B200_NET_FLAGS (e.g. -DJDN_PROFILE) and JITCDDE_B200_NET_BLOCK"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
from jitcdde_b200 import _build, _codegen
os.environ["JITCDDE_B200_CACHE"] = _build.PREBUILT
import sympy
program = _codegen.generate_network(lambda y, t, w: w, lambda yd, y: sympy.sin(yd-y), 1)
print(_build.build_net_image(program, mode=_build.FAST, verbose="-v" in sys.argv)[:0] or "built")
