"""
Builds and loads the host twin of the engine (tests/host_twin/twin.cpp) — TEST SCAFFOLDING.

The twin is the engine header compiled with g++ behind the same C ABI, so that the CPU
test-suite can run the product's Python classes and step logic without a GPU and compare
them with the oracle. It is selected only by the `use_twin` context manager below, which
patches jitcdde_b200._capi.load_library for the duration of a test.
"""

import contextlib
import ctypes
import os
import subprocess

from jitcdde_b200 import _build, _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
_OUT = os.path.join(_HERE, "host_twin", "_build")


def build_twin(program, n_basic=None):
	n_basic = n_basic or program.n
	window = int(_build.use_window(program))
	window = window if window < 2 else 0      # modes 2 and 3 (cp.async into shared memory) exist on the device only
	os.makedirs(_OUT, exist_ok=True)
	rhs = os.path.join(_OUT, f"rhs_{program.key}.inc")
	with open(rhs, "w") as f:
		f.write(program.body)
	so = os.path.join(_OUT, f"twin_{program.key}_{n_basic}_w{window}.so")
	src = os.path.join(_HERE, "host_twin", "twin.cpp")
	deps = [src] + [os.path.join(_build.CSRC, x) for x in ("jdde_engine.cuh", "jdde_team.cuh", "jdde_group.cuh", "jdde_math.h", "jdde_abi.h")]
	if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
		subprocess.run([
			"g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
			f"-DJD_N={program.n}", f"-DJD_NBASIC={n_basic}", f"-DJD_NSITES={program.n_sites}",
			f"-DJD_NPARAMS={program.n_params}", f'-DJD_RHS_FILE="{rhs}"', f"-DJD_WINDOW={window}",
			f"-I{_build.CSRC}", src, "-o", so, "-lm",
		], check=True)
	return _load(so)


def _load(so):
	lib = ctypes.CDLL(so)
	for name, (restype, argtypes) in _capi.SIGNATURES.items():
		if hasattr(lib, name):
			fn = getattr(lib, name)
			fn.restype = restype
			fn.argtypes = argtypes
	return lib


def build_net_twin(program):
	"""host twin of the network engine (tests/host_twin/net_twin.cpp) for one (local, coupling) pair"""
	os.makedirs(_OUT, exist_ok=True)
	rhs = os.path.join(_OUT, f"net_{program.key}.inc")
	with open(rhs, "w") as f:
		f.write(program.body)
	so = os.path.join(_OUT, f"net_twin_{program.key}.so")
	src = os.path.join(_HERE, "host_twin", "net_twin.cpp")
	deps = [src] + [os.path.join(_build.CSRC, x) for x in ("jdn_engine.cuh", "jdn_abi.h", "jdde_math.h", "jdde_abi.h")]
	if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
		subprocess.run([
			"g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
			f'-DJDN_RHS_FILE="{rhs}"', f"-I{_build.CSRC}", src, "-o", so, "-lm",
		], check=True)
	return _load(so)


class _Both:
	"""routes jdde_net_* to the network twin and everything else to the ensemble twin"""
	def __init__(self, state):
		self._state = state

	def __getattr__(self, name):
		return getattr(self._state["net_lib" if name.startswith("jdde_net_") else "lib"], name)


@contextlib.contextmanager
def use_twin():
	"""
	Within this context jitcdde_b200 talks to the host twin instead of libjitcdde_b200.so:
	`_build.build_image` compiles the twin for the generated program with g++ (instead of a
	cubin with nvcc), `_capi.load_library` returns it and `Engine.load_image` is a no-op.
	"""
	state = {}

	def fake_build_image(program, n_basic=None, **kwargs):
		state["lib"] = build_twin(program, n_basic)
		return "<host twin>"

	def fake_build_net_image(program, **kwargs):
		state["net_lib"] = build_net_twin(program)
		return "<host twin>"

	saved = _build.build_image, _build.build_net_image, _capi.load_library, _capi.Engine.load_image, _capi.NetEngine.load_image
	_build.build_image = fake_build_image
	_build.build_net_image = fake_build_net_image
	both = _Both(state)
	_capi.load_library = lambda path=None: both
	_capi.Engine.load_image = lambda self, path: None
	_capi.NetEngine.load_image = lambda self, path: None
	try:
		yield state
	finally:
		_build.build_image, _build.build_net_image, _capi.load_library, _capi.Engine.load_image, _capi.NetEngine.load_image = saved
