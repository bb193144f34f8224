"""
build_ref.py — compile the REFERENCE'S OWN C core into oracle/_ref/. TEST INFRASTRUCTURE ONLY.

Recipe: the unmodified template
/root/reference/jitcdde/jitced_template.c is rendered with Jinja2 using exactly the
variables the reference passes (/root/reference/jitcdde/_jitcdde.py:562-573 plus
`module_name`), the right-hand side is written next to it as `f.c` /
`f_definitions.c` (normally printed by SymEngine, which is not installed here) and
the result is compiled with gcc against Python.h + NumPy as a CPython extension
module, which is what the reference's `_compile_and_load` does through setuptools.

Nothing of the reference is copied into the repository: rendered sources and
binaries go to oracle/_ref/ (git-ignored; the built .so files travel to the GPU
box with the snapshot, where /root/reference does not exist).

Two flavours of the right-hand side:
  chain   – the statements of jitcdde_b200._codegen (common subexpressions merged,
            integer powers as a multiplication chain) bound to the reference's
            macros; bit-comparable with the CUDA verification build.
  literal – one `set_dy(i, …)` per component, nested `anchors(…)` calls, libm
            `pow`, i.e. what the reference itself would have generated.
"""

import importlib.util
import os
import subprocess
import sys
import sysconfig

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
TEMPLATE = "/root/reference/jitcdde/jitced_template.c"
_CSRC = os.path.join(os.path.dirname(_HERE), "jitcdde_b200", "csrc")

VERIFY_FLAGS = ["-O2", "-ffp-contract=off", "-g", "-UNDEBUG"]
FAST_FLAGS = ["-Ofast", "-DNDEBUG"]   # reference-like default flags (jitcxde_common: -Ofast), without -march=native: must run on the GPU box, too


def reference_available():
	return os.path.exists(TEMPLATE)


def _so_path(name):
	return os.path.join(REF_DIR, name + sysconfig.get_config_var("EXT_SUFFIX"))


def build_ref(program, name, n_basic=None, flavour="chain", flags=VERIFY_FLAGS, force=False, tangent_indices=(), definitions=""):
	"""Render + compile; returns the path of the extension module."""
	so = _so_path(name)
	if os.path.exists(so) and not force:
		return so
	if not reference_available():
		raise FileNotFoundError(f"{TEMPLATE} not present and {so} not prebuilt")
	import jinja2

	n_basic = n_basic or program.n
	work = os.path.join(REF_DIR, "src_" + name)
	os.makedirs(work, exist_ok=True)

	if flavour == "chain":
		sites = program.n_sites
		f_c = program.body
		first_par = f"self->parameter_{program.param_names[0]}" if program.param_names else "0"
		defs = f"""
#include "{os.path.join(_CSRC, 'jdde_math.h')}"
#define JDDE_Y(i) current_y(i)
#define JDDE_PAR(k) ((&{first_par})[k])
#define JDDE_SEG anchor * const
#define JDDE_ANCHORS(k, tm) anchors(tm)
#define JDDE_PAST_Y(tm, i, s) past_y(tm, i, s)
#define JDDE_PAST_DY(tm, i, s) past_dy(tm, i, s)
#define JDDE_SET_DY(i, value) set_dy(i, (value))
#define JDDE_IPOW(x, e) jdde_ipow(x, e)
"""
	elif flavour == "literal":
		sites = program.literal_sites
		f_c = "\n".join(program.literal_f) + "\n"
		defs = definitions      # e.g. an #include of csrc/jdde_math.h for right-hand sides written with its functions
	else:
		raise ValueError(flavour)

	with open(TEMPLATE) as f:
		template = f.read()
	code = jinja2.Environment().from_string(template).render(
		n=program.n,
		number_of_helpers=0,
		number_of_anchor_helpers=0,
		has_any_helpers=0,
		anchor_mem_length=sites,
		n_basic=n_basic,
		control_pars=list(program.param_names),
		tangent_indices=list(tangent_indices),
		chunk_size=100,
		callbacks=[],
		module_name=name,
	)
	with open(os.path.join(work, name + ".c"), "w") as f:
		f.write(code)
	with open(os.path.join(work, "f.c"), "w") as f:
		f.write(f_c)
	with open(os.path.join(work, "f_definitions.c"), "w") as f:
		f.write(defs)

	cmd = [
		"gcc", "-std=c11", *flags, "-fPIC", "-shared", "-Wno-unknown-pragmas",
		f"-I{sysconfig.get_paths()['include']}", f"-I{np.get_include()}",
		os.path.join(work, name + ".c"), "-o", so, "-lm",
	]
	subprocess.run(cmd, check=True, cwd=work)
	return so


def load_ref(name):
	"""Import a previously built reference module (works on the GPU box, too)."""
	so = _so_path(name)
	if not os.path.exists(so):
		raise FileNotFoundError(so)
	if name in sys.modules:
		return sys.modules[name]
	spec = importlib.util.spec_from_file_location(name, so)
	module = importlib.util.module_from_spec(spec)
	spec.loader.exec_module(module)
	sys.modules[name] = module
	return module


def have_ref(name):
	return os.path.exists(_so_path(name))
