"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol that
include/jitcdde_b200.h declares; the ctypes binding covers all of them. No compute calls here."""

import os
import re

import pytest

from jitcdde_b200 import _build, _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
	with open(os.path.join(ROOT, "include", "jitcdde_b200.h")) as f:
		text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
	return sorted(set(re.findall(r"\b(jdde_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
	lib = _capi.load_library(_build.build_runtime())
	names = declared_symbols()
	assert len(names) >= 30
	for name in names:
		assert hasattr(lib, name), f"libjitcdde_b200.so lacks {name}"
	assert set(names) == set(_capi.SIGNATURES), set(names) ^ set(_capi.SIGNATURES)


def test_no_cpu_fallback():
	"""Without a CUDA device the product fails loudly instead of computing on the CPU."""
	import torch
	if torch.cuda.is_available():
		pytest.skip("a GPU is present")
	with pytest.raises(_capi.EngineError, match="no CUDA device"):
		_capi.Engine(1, 1, 1, 0, 4)


def test_kernel_image_builds_for_sm100a():
	from tests import systems
	prog = systems.program(systems.mackey_glass())
	for mode in (_build.VERIFY, _build.FAST):
		path = _build.build_image(prog, mode=mode)
		with open(path, "rb") as f:
			assert f.read(4) == b"\x7fELF"


def test_product_does_not_import_oracle():
	pkg = os.path.join(ROOT, "jitcdde_b200")
	for directory, _, files in os.walk(pkg):
		for name in files:
			if name.endswith((".py", ".cu", ".cuh", ".h")):
				with open(os.path.join(directory, name)) as f:
					text = f.read()
				assert "import oracle" not in text and "from oracle" not in text and "dde_oracle" not in text.replace("oracle/dde_oracle.c", ""), name
