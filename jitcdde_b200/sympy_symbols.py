"""
SymPy flavour of the symbols (reference: /root/reference/jitcdde/sympy_symbols.py). The symbols of this package are
SymPy objects to begin with (SymEngine is not a dependency), so this module only re-exports them under the name the
reference's users import: `from jitcdde.sympy_symbols import t, y`.
"""

from ._symbols import anchors, current_y, dy, input, past_dy, past_y, quadrature, t, y  # noqa: A004,F401
