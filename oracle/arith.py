"""Number systems for the oracle (TEST INFRASTRUCTURE — see oracle/__init__.py).

`Dec15` restates the arithmetic the reference's hot path performs: every call site passes
`rounder = MathContext(15, RoundingMode.HALF_UP)` (LPState.java:18 and the `divide(…, rounder)`
`multiply(…, rounder)` `subtract(…, rounder)` `add(…, rounder)` calls at LPState.java:139-177,
297).  `compareTo` (LPState.java:278,294,299) is an exact comparison without rounding.
Inputs are *not* rounded on entry — BigDecimal keeps every parsed digit and the first
operation rounds its result.

`F64` is the binary64 twin: one IEEE round-to-nearest-even per operation, no fused
multiply-add (CPython floats never contract), true division.
"""
from __future__ import annotations

import decimal
from decimal import Decimal

_CTX = decimal.Context(
    prec=15,
    rounding=decimal.ROUND_HALF_UP,
    Emax=decimal.MAX_EMAX,
    Emin=decimal.MIN_EMIN,
    capitals=1,
    clamp=0,
    flags=[],
    traps=[decimal.InvalidOperation, decimal.DivisionByZero, decimal.Overflow],
)


class Dec15:
    """BigDecimal + MathContext(15, HALF_UP)."""

    name = "dec15"
    ZERO = Decimal(0)
    ONE = Decimal(1)
    DEF_EPSILON = Decimal(1).scaleb(-9)   # new BigDecimal(BigInteger.ONE, 9)   LPState.java:20
    DEF_INF = Decimal(1).scaleb(50)       # new BigDecimal(BigInteger.ONE, -50) LPState.java:21

    @staticmethod
    def conv(x):
        if isinstance(x, Decimal):
            return x
        if isinstance(x, (int, str)):
            return Decimal(x)
        if isinstance(x, float):
            return Decimal(x)  # exact expansion of the binary64 value
        return Decimal(float(x))

    @staticmethod
    def mul(a, b):
        return _CTX.multiply(a, b)

    @staticmethod
    def sub(a, b):
        return _CTX.subtract(a, b)

    @staticmethod
    def add(a, b):
        return _CTX.add(a, b)

    @staticmethod
    def div(a, b):
        return _CTX.divide(a, b)

    @staticmethod
    def neg(a):
        # BigDecimal.negate() without a MathContext: exact, and 0.negate() is 0 (no signed zero)
        return -a if a else abs(a)

    @staticmethod
    def abs(a):
        return abs(a)

    @staticmethod
    def cmp(a, b):
        return (a > b) - (a < b)

    @staticmethod
    def set_scale6(v):
        """BigDecimal.setScale(6, HALF_UP) (LPSolver.java:113)."""
        return v.quantize(Decimal("0.000001"), rounding=decimal.ROUND_HALF_UP)

    @staticmethod
    def to_float(a):
        return float(a)


class F64:
    """IEEE binary64, RN-even, separate mul/sub roundings, true division."""

    name = "f64"
    ZERO = 0.0
    ONE = 1.0
    DEF_EPSILON = 1e-9
    DEF_INF = 1e50

    @staticmethod
    def conv(x):
        return float(x)

    @staticmethod
    def mul(a, b):
        return a * b

    @staticmethod
    def sub(a, b):
        return a - b

    @staticmethod
    def add(a, b):
        return a + b

    @staticmethod
    def div(a, b):
        return a / b

    @staticmethod
    def neg(a):
        return -a

    @staticmethod
    def abs(a):
        return abs(a)

    @staticmethod
    def cmp(a, b):
        return (a > b) - (a < b)

    @staticmethod
    def set_scale6(v):
        return Decimal(v).quantize(Decimal("0.000001"), rounding=decimal.ROUND_HALF_UP)

    @staticmethod
    def to_float(a):
        return a
