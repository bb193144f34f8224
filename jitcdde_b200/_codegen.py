"""
Code generation: SymPy expressions of a DDE's right-hand side → one block of
C statements in a small macro vocabulary.

This replaces the reference's `compile_C` code path
(/root/reference/jitcdde/_jitcdde.py:420-575), which prints one
`set_dy(i, …)` statement per component through SymEngine's C printer into `f.c`
/ `helpers.c` and binds them to the C core with the macros at
/root/reference/jitcdde/jitced_template.c:382-391.

Here the *same token stream* is consumed by three back-ends, each of which only
defines the macros differently:

  * the CUDA engine (`csrc/jdde_engine.cuh`) — the product,
  * the plain-C restatement in `oracle/dde_oracle.c`              — test oracle,
  * the reference's own `jitced_template.c`, compiled unmodified
    by `oracle/build_ref.py` into `oracle/_ref/`                  — test oracle.

Feeding all three from one printer keeps the association order of every
floating-point expression identical, which is what makes bit-for-bit comparison
of step sequences possible (SURVEY.md §7 "hard parts" 1).

Vocabulary (all macros):
  JDDE_Y(i)                 current state component            (reference: current_y)
  JDDE_PAR(k)               k-th control parameter             (reference: self->parameter_<name>)
  JDDE_SEG                  type of an anchor-pair handle      (reference: anchor *)
  JDDE_ANCHORS(k, time)     bracket search at lookup site k    (reference: anchors → get_past_anchors)
  JDDE_PAST_Y(time, i, s)   cubic Hermite value                (reference: past_y  → get_past_value)
  JDDE_PAST_DY(time, i, s)  cubic Hermite derivative           (reference: past_dy → get_past_diff)
  JDDE_SET_DY(i, value)     store derivative component         (reference: set_dy)
  JDDE_IPOW(x, e)           x**e for a literal integer e as a fixed multiplication chain
  JDDE_SIN/COS/EXP/TANH/LOG elementary functions: libm in the production build, the bit-reproducible ones of
                            csrc/jdde_math.h in the verification build, oracle and "chain" reference core

Lookup sites: the reference gives every textual `anchors(…)` call its own
search cursor (`anchor_mem`, jitced_template.c:44-45, 193-217;
`past_calls`, _jitcdde.py:508-517). Two call sites with the same argument
expression see the same sequence of times from the same initial cursor, so
their cursors and results are always identical; merging them (which common
subexpression elimination does here) changes no result.
"""

from __future__ import annotations

import hashlib
from dataclasses import dataclass, field

import sympy
from sympy.printing.c import C99CodePrinter

from ._symbols import anchors, current_y, past_dy, past_y, t

_MAX_IPOW = 64

#: elementary functions that csrc/jdde_math.h implements with the same bits on CPU and GPU for the verification build
#: (the production build and the "literal" reference flavour call libm)
_ELEMENTARY = {"sin": "JDDE_SIN", "cos": "JDDE_COS", "exp": "JDDE_EXP", "tanh": "JDDE_TANH", "log": "JDDE_LOG"}


class _MacroPrinter(C99CodePrinter):
	"""Prints SymPy expressions in the JDDE_* macro vocabulary."""

	def __init__(self, par_index, seg_symbols, literal_pow=False):
		super().__init__({"allow_unknown_functions": True, "user_functions": dict(_ELEMENTARY)})
		self.par_index = par_index
		self.seg_symbols = seg_symbols
		self.literal_pow = literal_pow
		self.n_sites = 0

	def _print_Symbol(self, expr):
		if expr in self.par_index:
			return f"JDDE_PAR({self.par_index[expr]})"
		return super()._print_Symbol(expr)

	def _print_AppliedUndef(self, expr):
		# SymPy's dispatcher skips the concrete class of undefined functions
		handler = getattr(self, "_print_" + type(expr).__name__, None)
		if handler is None:
			return self._print_Function(expr)
		return handler(expr)

	def _print_current_y(self, expr):
		return f"JDDE_Y({self._print(expr.args[0])})"

	def _print_past_y(self, expr):
		tm, i, seg = expr.args
		return f"JDDE_PAST_Y({self._print(tm)}, {self._print(i)}, {self._print(seg)})"

	def _print_past_dy(self, expr):
		tm, i, seg = expr.args
		return f"JDDE_PAST_DY({self._print(tm)}, {self._print(i)}, {self._print(seg)})"

	def _print_anchors(self, expr):
		site = self.n_sites
		self.n_sites += 1
		return f"JDDE_ANCHORS({site}, {self._print(expr.args[0])})"

	def _print_Pow(self, expr):
		base, exp = expr.args
		if not self.literal_pow and exp.is_Integer and 0 < abs(int(exp)) <= _MAX_IPOW and int(exp) != 1:
			e = int(exp)
			core = f"JDDE_IPOW({self._print(base)}, {abs(e)})"
			return core if e > 0 else f"(1.0/{core})"
		return super()._print_Pow(expr)


@dataclass
class RHSProgram:
	"""Result of code generation for one DDE."""
	n: int
	param_names: list
	n_sites: int
	body: str                      # statements in the JDDE_* vocabulary
	uses_past_dy: bool
	literal_f: list = field(default_factory=list)   # one reference-style `set_dy(i, …)` per component, no CSE, libm pow
	literal_sites: int = 0
	group_lanes: int = 0           # blockwise Lyapunov program: 1 + n_lyap lanes can share a member (csrc/jdde_group.cuh); 0: not blockwise

	@property
	def n_params(self):
		return len(self.param_names)

	@property
	def key(self):
		h = hashlib.sha256()
		h.update(repr((self.n, self.param_names, self.n_sites)).encode())
		h.update(self.body.encode())
		return h.hexdigest()[:16]


def _substitute_helpers(exprs, helpers):
	"""Inline helpers (symbol, expression) into the expressions. Helpers may
	depend on earlier helpers; substitute until no helper symbol is left
	(the reference sorts them topologically, _jitcdde.py:207)."""
	if not helpers:
		return exprs
	subs = {sympy.sympify(s): sympy.sympify(e) for s, e in helpers}
	symbols = set(subs)
	out = []
	for expr in exprs:
		for _ in range(len(subs)+1):
			if not (expr.free_symbols & symbols):
				break
			expr = expr.xreplace(subs)
		else:
			raise ValueError("Helpers depend on each other cyclically.")
		out.append(expr)
	return out


def _to_plain_functions(expr):
	"""Map symbols that came from another package's flavour (e.g. a user's own
	`sympy.Function("current_y")`) onto ours by name."""
	ours = {f.__name__: f for f in (current_y, past_y, past_dy, anchors)}
	repl = {}
	for f in expr.atoms(sympy.Function):
		name = type(f).__name__
		if name in ours and type(f) is not ours[name]:
			repl[f] = ours[name](*f.args)
	for s in expr.free_symbols:
		if s.name == "t" and s != t:
			repl[s] = t
	return expr.xreplace(repl) if repl else expr


def generate(f_sym, helpers=(), control_pars=(), simplify=False, do_cse=True, blocks=None):
	"""
	f_sym: list of expressions (component i = dy_i/dt), helpers: list of
	(symbol, expression) incl. anchor helpers, control_pars: list of symbols.
	blocks = (n_basic, n_lyap): the system is a basic one followed by n_lyap tangent blocks of the same
	shape (tangent_vector_f); the statements are then generated block by block (see _generate_blocks), if
	the blocks turn out to be index-shifted copies of each other.
	"""
	control_pars = [sympy.sympify(p) for p in control_pars]
	exprs = [_to_plain_functions(sympy.sympify(e)) for e in f_sym]
	helpers = [(sympy.sympify(s), _to_plain_functions(sympy.sympify(e))) for s, e in helpers]
	exprs = _substitute_helpers(exprs, helpers)
	if simplify:
		exprs = [sympy.simplify(e, ratio=1.0) for e in exprs]
	n = len(exprs)

	allowed = {t, *control_pars}
	for i, e in enumerate(exprs):
		bad = e.free_symbols - allowed
		if bad:
			raise ValueError(f"Invalid symbol ({sorted(s.name for s in bad)}) in equation {i}.")

	par_index = {p: k for k, p in enumerate(control_pars)}

	# literal flavour: what the reference would emit (no CSE, libm pow)
	lit = _LiteralPrinter(par_index)
	literal_f = [f"set_dy({i}, {lit.doprint(e)});" for i, e in enumerate(exprs)]

	if blocks is not None and blocks[1] > 0:
		blockwise = _generate_blocks(exprs, blocks[0], blocks[1], par_index, do_cse)
		if blockwise is not None:
			body, n_sites = blockwise
			return RHSProgram(
				n=n,
				param_names=[p.name for p in control_pars],
				n_sites=n_sites,
				body=body,
				uses_past_dy=any(e.has(past_dy) for e in exprs),
				literal_f=literal_f,
				literal_sites=lit.n_sites,
				group_lanes=blocks[1]+1,
			)

	if do_cse:
		replacements, reduced = sympy.cse(exprs, symbols=sympy.numbered_symbols("x"), order="none")
	else:
		replacements, reduced = [], list(exprs)
	# every distinct anchors(...) must become a named temporary so that it is
	# evaluated exactly once per RHS evaluation
	replacements, reduced = _hoist_anchors(replacements, reduced)

	seg_symbols = {s for s, e in replacements if e.func == anchors}
	printer = _MacroPrinter(par_index, seg_symbols)
	lines = []
	for s, e in replacements:
		if e.func == anchors:
			lines.append(f"JDDE_SEG {s.name} = {printer.doprint(e)};")
		else:
			lines.append(f"const double {s.name} = {printer.doprint(e)};")
	for i, e in enumerate(reduced):
		lines.append(f"JDDE_SET_DY({i}, {printer.doprint(e)});")
	body = "\n".join(lines) + "\n"

	uses_past_dy = any(e.has(past_dy) for e in exprs)
	return RHSProgram(
		n=n,
		param_names=[p.name for p in control_pars],
		n_sites=printer.n_sites,
		body=body,
		uses_past_dy=uses_past_dy,
		literal_f=literal_f,
		literal_sites=lit.n_sites,
	)


def _hoist_anchors(replacements, reduced):
	"""Make every anchors(arg) call a statement of its own (`y0 = anchors(arg)`),
	placed right before its first use, and merge calls with identical argument."""
	out = []
	seg_of = {}          # anchors(arg) expression -> symbol
	counter = [0]

	def hoist(expr):
		calls = sorted(expr.atoms(anchors), key=sympy.default_sort_key)
		# inner-most first (nested delays)
		changed = True
		while changed:
			changed = False
			for call in sorted(expr.atoms(anchors), key=lambda c: c.count_ops()):
				arg = call.args[0]
				if arg.has(anchors):
					continue
				if call not in seg_of:
					sym = sympy.Symbol(f"s{counter[0]}")
					counter[0] += 1
					seg_of[call] = sym
					out.append((sym, call))
				expr = expr.xreplace({call: seg_of[call]})
				changed = True
				break
		del calls
		return expr

	for s, e in replacements:
		if e.func == anchors and not e.args[0].has(anchors):
			if e in seg_of:
				# same lookup already hoisted under another name: alias it
				out.append((s, seg_of[e]))
			else:
				seg_of[e] = s
				out.append((s, e))
			continue
		out.append((s, hoist(e)))
	new_reduced = [hoist(e) for e in reduced]
	# aliases `s = other_symbol` are removed by substitution
	alias = {s: e for s, e in out if isinstance(e, sympy.Symbol)}
	if alias:
		out = [(s, e.xreplace(alias)) for s, e in out if s not in alias]
		new_reduced = [e.xreplace(alias) for e in new_reduced]
	return out, new_reduced


# ---------------------------------------------------------------------------
# Blockwise programs for Lyapunov systems
# ---------------------------------------------------------------------------

class _NotBlockwise(Exception):
	pass


class _BlockPrinter(_MacroPrinter):
	"""Prints the statements of one block: components below n_basic in the usual vocabulary, components of
	the block itself (offset … offset+n_basic−1) relative to it as JDDE_TY / JDDE_PAST_TY / JDDE_PAST_TDY."""

	def __init__(self, par_index, n_basic, offset):
		super().__init__(par_index, set())
		self.n_basic = n_basic
		self.offset = offset

	def _split(self, index):
		if not index.is_Integer:
			raise _NotBlockwise
		index = int(index)
		if 0 <= index < self.n_basic:
			return False, index
		if self.offset is not None and self.offset <= index < self.offset+self.n_basic:
			return True, index-self.offset
		raise _NotBlockwise

	def _print_current_y(self, expr):
		own, index = self._split(expr.args[0])
		return f"JDDE_TY({index})" if own else f"JDDE_Y({index})"

	def _print_past_y(self, expr):
		tm, i, seg = expr.args
		own, index = self._split(i)
		return f"JDDE_PAST_{'TY' if own else 'Y'}({self._print(tm)}, {index}, {self._print(seg)})"

	def _print_past_dy(self, expr):
		tm, i, seg = expr.args
		own, index = self._split(i)
		return f"JDDE_PAST_{'TDY' if own else 'DY'}({self._print(tm)}, {index}, {self._print(seg)})"

	def _print_anchors(self, expr):
		raise _NotBlockwise      # all lookups are made in the prologue


def _extract_lookups(exprs):
	"""Replace every anchors(argument) call by a symbol s_k (inner-most first; calls with the same argument
	share one). Returns ([(s_k, argument)], expressions); arguments may contain earlier s_j (nested delays)."""
	seg_of = {}
	lookups = []
	out = []
	for expr in exprs:
		while True:
			calls = [c for c in expr.atoms(anchors) if not c.args[0].has(anchors)]
			if not calls:
				break
			for call in sorted(calls, key=sympy.default_sort_key):
				if call not in seg_of:
					seg_of[call] = sympy.Symbol(f"s{len(seg_of)}")
					lookups.append((seg_of[call], call.args[0]))
			expr = expr.xreplace({c: seg_of[c] for c in calls})
		out.append(expr)
	return lookups, out


def _generate_blocks(exprs, n_basic, n_lyap, par_index, do_cse):
	"""
	Statements for a basic system followed by n_lyap tangent blocks
	(/root/reference/jitcdde/_jitcdde.py:1046-1089). The tangent blocks are the same expressions with the block's
	own components shifted by n_basic each; they are printed ONCE, relative to the block, and

	  * the one-thread-per-member engine, the plain-C oracle and the reference's template get them n_lyap times
	    with JDDE_TOFF = n_basic, 2·n_basic, … (the #ifndef JDDE_GROUP branch), while
	  * the lane-group engine (csrc/jdde_group.cuh) gives the basic system to one lane and one tangent block
	    to each further lane of a group (the #ifdef JDDE_GROUP branch).

	All lookups are made first, by everyone. Both branches consist of the same statements, so the arithmetic
	per component is identical. Returns (body, n_sites), or None if the blocks are not index-shifted copies of
	each other (then the caller generates the system as a whole).
	"""
	if len(exprs) != n_basic*(n_lyap+1):
		return None
	lookups, exprs = _extract_lookups(exprs)

	def statements(block_exprs, printer, temp_prefix, setter):
		if do_cse:
			replacements, reduced = sympy.cse(block_exprs, symbols=sympy.numbered_symbols(temp_prefix), order="none")
		else:
			replacements, reduced = [], list(block_exprs)
		lines = [f"const double {s.name} = {printer.doprint(e)};" for s, e in replacements]
		lines += [f"{setter}({i}, {printer.doprint(e)});" for i, e in enumerate(reduced)]
		return "\n".join(lines) + "\n"

	try:
		basic_printer = _BlockPrinter(par_index, n_basic, None)
		prologue = ""
		for k, (s, argument) in enumerate(lookups):
			prologue += f"JDDE_SEG {s.name} = JDDE_ANCHORS({k}, {basic_printer.doprint(argument)});\n"
		basic = statements(exprs[:n_basic], basic_printer, "x", "JDDE_SET_DY")
		tangent = None
		for i in range(n_lyap):
			offset = (i+1)*n_basic
			text = statements(exprs[offset:offset+n_basic], _BlockPrinter(par_index, n_basic, offset), "w", "JDDE_SET_TDY")
			if tangent is None:
				tangent = text
			elif text != tangent:
				return None
	except _NotBlockwise:
		return None

	body = (
		"#ifndef JDDE_GROUP\n"
		"#define JDDE_TY(k) JDDE_Y((k)+JDDE_TOFF)\n"
		"#define JDDE_PAST_TY(tm, k, s) JDDE_PAST_Y(tm, (k)+JDDE_TOFF, s)\n"
		"#define JDDE_PAST_TDY(tm, k, s) JDDE_PAST_DY(tm, (k)+JDDE_TOFF, s)\n"
		"#define JDDE_SET_TDY(c, value) JDDE_SET_DY((c)+JDDE_TOFF, value)\n"
		"#endif\n"
		+ prologue +
		"#ifndef JDDE_GROUP\n"
		"{\n" + basic + "}\n"
	)
	for i in range(n_lyap):
		body += f"#define JDDE_TOFF {(i+1)*n_basic}\n{{\n{tangent}}}\n#undef JDDE_TOFF\n"
	body += (
		"#undef JDDE_TY\n#undef JDDE_PAST_TY\n#undef JDDE_PAST_TDY\n#undef JDDE_SET_TDY\n"
		"#else\n"
		"if (JDDE_GROUP_IS_BASIC)\n{\n" + basic + "}\nelse\n{\n" + tangent + "}\n"
		"#endif\n"
	)
	return body, len(lookups)


class _LiteralPrinter(C99CodePrinter):
	"""Reference vocabulary (jitced_template.c:382-391), no CSE, `pow` for all powers."""

	def __init__(self, par_index):
		super().__init__({"allow_unknown_functions": True})
		self.par_names = {p: f"self->parameter_{p.name}" for p in par_index}
		self.n_sites = 0

	def _print_Symbol(self, expr):
		if expr in self.par_names:
			return self.par_names[expr]
		return super()._print_Symbol(expr)

	def _print_AppliedUndef(self, expr):
		if type(expr).__name__ == "anchors":
			return self._print_anchors(expr)
		return self._print_Function(expr)

	def _print_anchors(self, expr):
		self.n_sites += 1
		return f"anchors({self._print(expr.args[0])})"


# ---------------------------------------------------------------------------
# Tangent-vector system for Lyapunov exponents
# ---------------------------------------------------------------------------

def get_delays(exprs, helpers=()):
	"""[0, *delays] with delay = t − argument of every anchors(...) call
	(/root/reference/jitcdde/_jitcdde.py:60-67)."""
	terms = set()
	for e in list(exprs) + [h[1] for h in helpers]:
		e = _to_plain_functions(sympy.sympify(e))
		terms.update(call.args[0] for call in e.atoms(anchors))
	delays = [sympy.simplify(t - term) for term in sorted(terms, key=sympy.default_sort_key)]
	return [0, *delays]


def jacobians(f_basic, n, delays, simplify=True):
	"""jacs[d][c][k] = ∂f_c/∂y_k(t − delays[d]) (/root/reference/jitcdde/_jitcdde.py:1046-1063); helpers already substituted"""
	from ._symbols import y
	jacs = []
	for delay in delays:
		rows = []
		for entry in f_basic:
			row = []
			for k in range(n):
				var = y(k, t - delay)
				dummy = sympy.Dummy()
				d = entry.xreplace({var: dummy}).diff(dummy).xreplace({dummy: var})
				if simplify:
					d = sympy.simplify(d)
				row.append(d)
			rows.append(row)
		jacs.append(rows)
	return jacs


def tangent_vector_f(f_basic, helpers, n, n_lyap, delays, simplify=True):
	"""
	Extended system (main dynamics + n_lyap tangent vectors), following
	/root/reference/jitcdde/_jitcdde.py:1046-1089: for exponent i and component
	c the derivative is Σ_delays Σ_k ∂f_c/∂y_k(t−δ) · y_{k+(i+1)n}(t−δ).
	"""
	from ._symbols import y
	f_basic = [_to_plain_functions(sympy.sympify(e)) for e in f_basic]
	helpers = [(sympy.sympify(s), _to_plain_functions(sympy.sympify(e))) for s, e in helpers]
	f_basic = _substitute_helpers(f_basic, helpers)

	jacs = jacobians(f_basic, n, delays, simplify=simplify)

	out = list(f_basic)
	for i in range(n_lyap):
		for c in range(n):
			expression = sympy.Integer(0)
			for delay, jac in zip(delays, jacs, strict=True):
				for k, entry in enumerate(jac[c]):
					expression += entry * y(k + (i+1)*n, t - delay)
			if simplify:
				expression = sympy.simplify(expression, ratio=1.0)
			out.append(expression)
	return out


# ---------------------------------------------------------------------------
# Network engine: local and coupling function of a delay-coupled network
# ---------------------------------------------------------------------------

@dataclass
class NetProgram:
	"""Generated macros JDN_NPARAMS / JDN_LOCAL / JDN_COUPLING for csrc/jdn_engine.cuh."""
	n_node_params: int
	body: str

	@property
	def key(self):
		return hashlib.sha256(self.body.encode()).hexdigest()[:16]


def generate_network(local, coupling, n_node_params=0):
	"""
	local(y, t, *node_parameters) and coupling(y_delayed, y) return SymPy expressions of their (symbolic)
	arguments:  dy_i/dt = local(y_i, t, p_i…) + Σ_e weight_e · coupling(y_src(t − tau_e), y_i).
	"""
	y_sym, yd_sym, t_sym = sympy.Symbol("jdn_y"), sympy.Symbol("jdn_yd"), sympy.Symbol("jdn_t")
	pars = [sympy.Symbol(f"jdn_p{k}") for k in range(n_node_params)]
	local_expr = sympy.sympify(local(y_sym, t_sym, *pars))
	coupling_expr = sympy.sympify(coupling(yd_sym, y_sym))
	for name, expr, allowed in (("local", local_expr, {y_sym, t_sym, *pars}), ("coupling", coupling_expr, {yd_sym, y_sym})):
		bad = expr.free_symbols - allowed
		if bad:
			raise ValueError(f"Invalid symbol ({sorted(s.name for s in bad)}) in the {name} function.")
	printer = _MacroPrinter({}, set())
	names = {y_sym: "(y)", yd_sym: "(yd)", t_sym: "(t)", **{p: f"(par[{k}])" for k, p in enumerate(pars)}}
	printer._print_Symbol = lambda expr: names.get(expr, expr.name)
	body = (
		f"#define JDN_NPARAMS {n_node_params}\n"
		f"#define JDN_LOCAL(y, t, par) ({printer.doprint(local_expr)})\n"
		f"#define JDN_COUPLING(yd, y) ({printer.doprint(coupling_expr)})\n"
	)
	return NetProgram(n_node_params=n_node_params, body=body)
