"""
Building the native parts for sm_100a: the runtime library (C ABI,
include/jitcdde_b200.h) and one kernel image (cubin) per right-hand side.

Counterpart of the reference's `_compile_and_load` via jitcxde_common/setuptools
(/root/reference/jitcdde/_jitcdde.py:560-575) and of `save_compiled` /
`module_location` (:178-179): images are cached in-tree under
jitcdde_b200/_cache/, keyed by a hash of the generated code, the engine sources
and the compiler flags, so a parameter sweep compiles once.

nvcc cross-compiles without a GPU. Everything is built for sm_100a only.
"""

import hashlib
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
INCLUDE = os.path.join(os.path.dirname(_PKG), "include")
PREBUILT = os.path.join(_PKG, "_cache")      # images that travel with the tree (read-mostly)


def _user_cache():
	explicit = os.environ.get("JITCDDE_B200_CACHE")
	if explicit:
		return explicit
	base = os.environ.get("XDG_CACHE_HOME") or os.path.join(os.path.expanduser("~"), ".cache")
	return os.path.join(base, "jitcdde_b200")


CACHE = _user_cache()
RUNTIME_LIB = os.path.join(_PKG, "libjitcdde_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

#: verification build: no FMA contraction → bit-identical with the CPU oracle (hard part 1)
VERIFY = "verify"
#: production build: FMA contraction allowed (nvcc default)
FAST = "fast"


def nvcc():
	path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
	if not os.path.exists(path):
		raise RuntimeError("nvcc not found; the engine has no CPU fallback and cannot be built without the CUDA toolkit")
	return path


def _sources_hash(sources, extra=""):
	h = hashlib.sha256(extra.encode())
	for path in sources:
		with open(path, "rb") as f:
			h.update(f.read())
	return h.hexdigest()


def cache_dir():
	"""where new images are written (created on demand)"""
	path = _user_cache()
	os.makedirs(path, exist_ok=True)
	return path


def _cached(name):
	"""path of a cached file if it exists (prebuilt directory first), marking it as recently used"""
	for root in (PREBUILT, _user_cache()):
		path = os.path.join(root, name)
		if os.path.exists(path):
			try:
				os.utime(path)
			except OSError:
				pass
			return path
	return None


def _evict(limit_mb=None):
	"""keep the user cache below the limit: least recently used images first"""
	limit = float(os.environ.get("JITCDDE_B200_CACHE_MB", limit_mb or 2048))*1e6
	root = _user_cache()
	try:
		entries = [(e.stat().st_mtime, e.stat().st_size, e.path) for e in os.scandir(root) if e.is_file()]
	except OSError:
		return
	total = sum(size for _, size, _ in entries)
	for _, size, path in sorted(entries):
		if total <= limit:
			break
		try:
			os.remove(path)
			total -= size
		except OSError:
			pass


def _write_atomically(path, text):
	"""generated include files: written under a process-private name and renamed (several ranks generate the same file)"""
	try:
		with open(path) as f:
			if f.read() == text:
				return
	except OSError:
		pass
	tmp = f"{path}.{os.getpid()}.tmp"
	with open(tmp, "w") as f:
		f.write(text)
	os.replace(tmp, path)


def _stale(target, stamp_value):
	"""A built file is current if its stamp file holds the hash of the sources it was built from (content, not
	mtimes: the tree is copied to other machines, and several ranks may import the package at once)."""
	try:
		with open(target + ".stamp") as f:
			return not os.path.exists(target) or f.read().strip() != stamp_value
	except OSError:
		return True


def build_runtime(force=False, verbose=False):
	"""nvcc -shared … csrc/jdde_runtime.cu csrc/jdn_runtime.cu → jitcdde_b200/libjitcdde_b200.so"""
	sources = [os.path.join(CSRC, "jdde_runtime.cu"), os.path.join(CSRC, "jdn_runtime.cu"), os.path.join(CSRC, "jdde_abi.h"),
		os.path.join(CSRC, "jdn_abi.h"), os.path.join(INCLUDE, "jitcdde_b200.h")]
	stamp = _sources_hash(sources)
	if force or _stale(RUNTIME_LIB, stamp):
		tmp = f"{RUNTIME_LIB}.{os.getpid()}.tmp"
		cmd = [nvcc(), "-shared", "-Xcompiler", "-fPIC", "-O2", "-std=c++17", *ARCH, "-lineinfo",
			sources[0], sources[1], "-o", tmp]
		if verbose:
			print(" ".join(cmd))
		subprocess.run(cmd, check=True)
		os.replace(tmp, RUNTIME_LIB)
		with open(RUNTIME_LIB + ".stamp", "w") as f:
			f.write(stamp)
	return RUNTIME_LIB


def _engine_hash():
	h = hashlib.sha256()
	for name in ("jdde_engine.cuh", "jdde_team.cuh", "jdde_group.cuh", "jdde_kernels.cu", "jdde_math.h", "jdde_abi.h"):
		with open(os.path.join(CSRC, name), "rb") as f:
			h.update(f.read())
	with open(os.path.join(INCLUDE, "jitcdde_b200.h"), "rb") as f:
		h.update(f.read())
	return h.hexdigest()[:12]


def _anchor_doubles(program):
	return ((1+2*program.n)+3) & ~3


def use_window(program):
	"""
	How lookups reach the anchors around their cursor (JD_WINDOW in csrc/jdde_engine.cuh):
	  0  speculative vector loads through L1/L2 at every lookup,
	  1  a register-resident window (measured slower: its registers cost resident warps),
	  2  a per-thread window in shared memory filled by cp.async one move ahead (fastest where it fits:
	     1.29e10 → 1.77e10 steps/s for the Mackey–Glass ensemble, profiles/variants_r01.md).
	  3  the same for anchors of more than 8 doubles: contiguous anchors in a per-thread area of shared memory,
	     lookups hand out pointers into it (n = 6 with one distinct delay: 520 bytes per thread).
	  4  an 8-anchor window (cursor−3 … cursor+4, 24 bytes per anchor for n = 1) with one cp.async group per anchor and
	     wait_group: no lookup waits for a copy any more, but more instructions per move; measured equal to mode 2
	     within 2 % (profiles/variants_r02.md), kept as a variant (JITCDDE_B200_WINDOW=4).
	Mode 2 needs 4 anchors × lookup sites × 8·A bytes of shared memory per thread; it is chosen when that is
	at most 128 bytes (n = 1 with one distinct delay), so that 32 warps per SM stay resident. Mode 3 is chosen up to
	600 bytes per thread (64-thread blocks, 6 per SM), unless the system is a Lyapunov extension that the lane-group
	stepper takes (use_group). JITCDDE_B200_WINDOW overrides (tests exercise all modes).
	"""
	A = _anchor_doubles(program)
	override = os.environ.get("JITCDDE_B200_WINDOW")
	if override is not None:
		mode = int(override) if program.n_sites > 0 else 0
		if mode == 3 and A <= 8:
			mode = 2
		if mode in (2, 4) and A > 8:
			mode = 3 if program.n_sites*4*A*8+8 <= 600 else 0
		return mode
	if program.n_sites > 0 and program.n_sites*4*A*8 <= 128:
		return 2
	if program.n_sites > 0 and A > 8 and program.n_sites*4*A*8+8 <= 600 and not getattr(program, "group_lanes", 0):
		return 3
	return 0


def launch_shape(program):
	"""
	(block size, resident blocks per SM) the kernels are compiled for. The integrate kernel is bound by
	latency (dependent FP64 chains and scattered anchor loads), not by issue slots or bandwidth, so resident
	warps matter more than spill-free code: small systems are compiled for one-warp blocks (a finished warp
	frees its slot at once) — 28 per SM (72 registers per thread) with the shared-memory window, 20 per SM
	(<= 96 registers) without — the measured optima on B200 (profiles/variants_r01.md, variants_r02.md). Larger
	systems keep all registers.
	"""
	override = os.environ.get("JITCDDE_B200_LAUNCH")
	if override:
		block, blocks = override.split("x")
		return int(block), int(blocks)
	if use_window(program) == 3:
		return 64, 6
	if program.n <= 2:
		return (32, 28) if use_window(program) in (2, 4) else (32, 20)
	if program.n <= 8:
		return 128, 4
	return 128, 1


def use_group(program):
	"""
	Lanes per member of the adaptive stepper (JD_GROUP in csrc/jdde_group.cuh), 0 = one thread per member.
	Blockwise Lyapunov programs (_codegen._generate_blocks) whose anchors are too large for registers anyway
	(more than 8 doubles: the by-pointer access path) are integrated by 1 + n_lyap lanes per member: one thread
	would need 16·n registers (local memory at n = 30), a lane needs 16·n_basic.
	JITCDDE_B200_GROUP=0 switches back to one thread per member (tests compare both).
	"""
	if os.environ.get("JITCDDE_B200_GROUP", "1") == "0":
		return 0
	lanes = getattr(program, "group_lanes", 0)
	if lanes and lanes <= 32 and _anchor_doubles(program) > 8 and use_window(program) == 0:
		return lanes
	return 0


def group_shape(program):
	"""(block size, resident blocks per SM) of jdde_k_integrate_group; JITCDDE_B200_GROUP_LAUNCH overrides"""
	override = os.environ.get("JITCDDE_B200_GROUP_LAUNCH")
	if override:
		block, blocks = override.split("x")
		return int(block), int(blocks)
	return 64, 8


def group_window(program):
	"""
	Shared-memory staging of the anchors a step can reach for the lane-group stepper (JD_GROUP_WINDOW in
	csrc/jdde_group.cuh): 4 whole anchors per member and lookup site. Chosen when that is at most 4 KB per member
	(two Rösslers + 4 exponents: 2 KB → 24 KB per block of 12 members); JITCDDE_B200_GROUP_WINDOW overrides.
	"""
	if not use_group(program):
		return 0
	override = os.environ.get("JITCDDE_B200_GROUP_WINDOW")
	if override is not None:
		return int(override) if program.n_sites > 0 else 0
	return int(program.n_sites > 0 and program.n_sites*4*_anchor_doubles(program)*8 <= 4096)


def image_path(program, n_basic, mode, block, extra_flags=()):
	tag = hashlib.sha256(repr((program.key, n_basic, mode, block, tuple(extra_flags), use_window(program), launch_shape(program), use_group(program), group_shape(program), group_window(program), _engine_hash())).encode()).hexdigest()[:20]
	return os.path.join(cache_dir(), f"jdde_{tag}.cubin")


def build_image(program, n_basic=None, mode=FAST, block=None, extra_flags=(), force=False, verbose=False):
	"""
	Compile the kernel image for one generated right-hand side
	(`program` = _codegen.RHSProgram). Returns the path of the cubin.
	"""
	n_basic = n_basic or program.n
	default_block, default_blocks = launch_shape(program)
	block = block or default_block
	if not any(f.startswith("-DJD_MIN_BLOCKS") for f in extra_flags):
		extra_flags = (*extra_flags, f"-DJD_MIN_BLOCKS={default_blocks}")
	path = image_path(program, n_basic, mode, block, extra_flags)
	hit = None if force else _cached(os.path.basename(path))
	if hit:
		return hit
	rhs = os.path.join(cache_dir(), f"rhs_{program.key}.inc")
	_write_atomically(rhs, program.body)
	flags = ["-O3", "-std=c++17", "-lineinfo", *ARCH]
	group_flags = []
	if use_group(program):
		group_block, group_blocks = group_shape(program)
		group_flags = [f"-DJD_GROUP={use_group(program)}", f"-DJD_GROUP_BLOCK={group_block}", f"-DJD_GROUP_MIN_BLOCKS={group_blocks}", f"-DJD_GROUP_WINDOW={group_window(program)}"]
	if mode == VERIFY:
		flags.append("-fmad=false")
	elif mode == FAST:
		flags.append("-DJD_FAST=1")
	else:
		raise ValueError(mode)
	tmp = f"{path}.{os.getpid()}.tmp"      # several ranks may build the same image at once: every process its own temporary, the rename is atomic
	cmd = [nvcc(), "-cubin", *flags,
		f"-DJD_N={program.n}", f"-DJD_NBASIC={n_basic}", f"-DJD_NSITES={program.n_sites}",
		f"-DJD_NPARAMS={program.n_params}", f'-DJD_RHS_FILE="{rhs}"', f"-DJD_BLOCK={block}",
		f"-DJD_WINDOW={int(use_window(program))}",
		*group_flags,
		*extra_flags,
		os.path.join(CSRC, "jdde_kernels.cu"), "-o", tmp]
	if verbose:
		print(" ".join(cmd))
	result = subprocess.run(cmd, capture_output=True, text=True)
	if result.returncode:
		raise RuntimeError(f"nvcc failed:\n{result.stderr}\ncommand: {' '.join(cmd)}")
	os.replace(tmp, path)
	_evict()
	return path


def build_net_image(program, block=256, mode=FAST, force=False, verbose=False):
	"""Kernel image of the network engine for one (local, coupling) pair (`program` = _codegen.NetProgram); `mode` as
	for build_image: VERIFY compiles without FMA contraction and with the bit-reproducible elementary functions."""
	h = hashlib.sha256()
	for name in ("jdn_engine.cuh", "jdn_kernels.cu", "jdn_abi.h", "jdde_abi.h", "jdde_math.h"):
		with open(os.path.join(CSRC, name), "rb") as f:
			h.update(f.read())
	extra = tuple(os.environ.get("JITCDDE_B200_NET_FLAGS", "").split())      # experiments: e.g. -DJDN_RUN_MIN_BLOCKS=6
	block = int(os.environ.get("JITCDDE_B200_NET_BLOCK", block))
	tag = hashlib.sha256(repr((program.key, block, mode, extra, h.hexdigest())).encode()).hexdigest()[:20]
	path = os.path.join(cache_dir(), f"jdn_{tag}.cubin")
	hit = None if force else _cached(os.path.basename(path))
	if hit:
		return hit
	rhs = os.path.join(cache_dir(), f"net_{program.key}.inc")
	_write_atomically(rhs, program.body)
	if mode == VERIFY:
		mode_flags = ["-fmad=false"]
	elif mode == FAST:
		mode_flags = ["-DJD_FAST=1"]
	else:
		raise ValueError(mode)
	tmp = f"{path}.{os.getpid()}.tmp"
	cmd = [nvcc(), "-cubin", "-O3", "-std=c++17", "-lineinfo", *ARCH, *mode_flags, *extra, f'-DJDN_RHS_FILE="{rhs}"', f"-DJDN_BLOCK={block}",
		os.path.join(CSRC, "jdn_kernels.cu"), "-o", tmp]
	if verbose:
		print(" ".join(cmd))
	result = subprocess.run(cmd, capture_output=True, text=True)
	if result.returncode:
		raise RuntimeError(f"nvcc failed:\n{result.stderr}\ncommand: {' '.join(cmd)}")
	os.replace(tmp, path)
	_evict()
	return path


def build_peak_image(force=False):
	"""cubin of the FP64 throughput microbenchmark (csrc/jdde_peak.cu)"""
	src = os.path.join(CSRC, "jdde_peak.cu")
	name = f"jdde_peak_{_sources_hash([src])[:16]}.cubin"
	hit = None if force else _cached(name)
	if hit:
		return hit
	path = os.path.join(cache_dir(), name)
	tmp = f"{path}.{os.getpid()}.tmp"
	subprocess.run([nvcc(), "-cubin", "-O3", "-lineinfo", *ARCH, src, "-o", tmp], check=True)
	os.replace(tmp, path)
	return path


# ---------------------------------------------------------------------------
# save_compiled / module_location (reference: jitcxde_common's save_compiled, _jitcdde.py:178-179, tests/test_jitcdde.py:138-147)

_CONTAINER_MAGIC = b"JDDEIMG1"


def save_image(path, image, meta, overwrite=False):
	"""one file: magic, length of the JSON header, header (what the engine needs to know about the system), kernel image"""
	import json
	if os.path.exists(path) and not overwrite:
		raise OSError(f"Target File {path} already exists and overwriting is disabled")
	header = json.dumps(meta).encode()
	with open(image, "rb") as f:
		blob = f.read()
	with open(path, "wb") as f:
		f.write(_CONTAINER_MAGIC + len(header).to_bytes(4, "little") + header + blob)
	return path


def load_image_file(path):
	"""(meta, path of the kernel image extracted into the cache) of a file written by save_image"""
	import json
	with open(path, "rb") as f:
		data = f.read()
	if data[:8] != _CONTAINER_MAGIC:
		raise ValueError(f"{path} is not a module file written by save_compiled")
	n = int.from_bytes(data[8:12], "little")
	meta = json.loads(data[12:12+n].decode())
	blob = data[12+n:]
	name = "loaded_" + hashlib.sha256(blob).hexdigest()[:20] + ".cubin"
	hit = _cached(name)
	if hit:
		return meta, hit
	target = os.path.join(cache_dir(), name)
	with open(f"{target}.{os.getpid()}.tmp", "wb") as f:
		f.write(blob)
	os.replace(f"{target}.{os.getpid()}.tmp", target)
	_evict()
	return meta, target
