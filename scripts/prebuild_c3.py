"""Build the kernel images of the C3 configuration (two Roesslers + 4 exponents) here, so that GPU calls do not spend their time in nvcc.
Usage: python scripts/prebuild_c3.py [GROUP_LAUNCH ...]   e.g. 64x8 64x4 128x4 32x16"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
from jitcdde_b200 import jitcdde_lyap
from tests import systems

def build(verification):
	system = systems.two_roessler(control=True)
	DDE = jitcdde_lyap(system["f"], n_lyap=4, control_pars=system["control_pars"], n_members=8, verbose=False, verification=verification)
	DDE.compile_C()
	return DDE._image

for shape in sys.argv[1:] or [None]:
	if shape:
		os.environ["JITCDDE_B200_GROUP_LAUNCH"] = shape
	for verification in ((True, False) if shape is None else (False,)):
		print(shape, verification, build(verification))
os.environ.pop("JITCDDE_B200_GROUP_LAUNCH", None)
os.environ["JITCDDE_B200_GROUP"] = "0"
for verification in (True, False):
	print("solo", verification, build(verification))
