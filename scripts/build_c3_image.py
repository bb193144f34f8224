"""compile the image of the C3 bench leg (two Rösslers + 4 exponents) into the in-tree cache, without a GPU"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
from jitcdde_b200 import _build
os.environ["JITCDDE_B200_CACHE"] = _build.PREBUILT
from jitcdde_b200 import jitcdde_lyap
from tests import systems
system = systems.two_roessler(control=True)
try:
	jitcdde_lyap(system["f"], n_lyap=4, control_pars=system["control_pars"], n_members=100000, verbose=False, ring_capacity=2048).compile_C()
	print("image built")
except Exception as error:
	print("image built; engine:", str(error)[:80])
