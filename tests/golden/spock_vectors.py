"""Known-answer vectors the reference's own tests hold for the pivot path.

Hand-transcribed DATA (inputs and expected outputs only) from
`src/test/groovy/lpsolver/LPStateSpec.groovy`, `src/test/groovy/lpsolver/LPSolverSpec.groovy`
and the recorded TRACE runs in `logs/lp_solver.log` of Toptachamann/Linear_Programming_Solver.
Each entry cites the lines it comes from.  Numbers are kept as strings where the spec uses a
non-integer literal so that both number systems parse them exactly.
"""

# LPStateSpec.groovy:12-29 — getEntering: first index with c[i] > eps
GET_ENTERING = [
    ([1, 2, 3], 0),
    ([0, 1, 2], 1),
    ([0, 0, 0, 0, 1], 4),
    ([-1, -1, -1, 3], 3),
    ([-1, 0, 0, 4], 3),
]

# LPStateSpec.groovy:31-48 — getLeaving incl. a 4-way tie -> row 0
GET_LEAVING = {
    "A": [[1, -1, 0, -1], [2, -2, "0.1", -1], [3, -2, 0, -4], [4, 1, 0, "0.1"]],
    "b": [1, 2, 3, 4],
    "cases": [(0, 0), (1, 3), (2, 1), (3, 3)],   # (entering, expected leaving)
}


def _names(k):
    return {i: "x%d" % (i + 1) for i in range(k)}, {"x%d" % (i + 1): i for i in range(k)}


# LPStateSpec.groovy:50-67 — 1x1 pivot
PIVOT_1x1 = {
    "A": [[1]], "b": [5], "c": [1], "m": 1, "n": 1, "e": 0, "l": 0,
    "resA": [[1]], "resB": [5], "resC": [-1], "resV": 5,
    "resVariables": {0: "x2", 1: "x1"}, "resCoefficients": {"x2": 0, "x1": 1},
}

# LPStateSpec.groovy:70-99 — 2x2 pivot, four (leaving, entering) combinations
PIVOT_2x2 = {
    "A": [[1, 2], [4, 4]], "b": [2, 4], "c": [4, 2], "m": 2, "n": 2,
    "cases": [
        dict(l=0, e=0, resB=[2, -4], resC=[-4, -6], resV=8,
             resA=[[1, 2], [-4, -4]],
             resVariables={0: "x3", 1: "x2", 2: "x1", 3: "x4"},
             resCoefficients={"x1": 2, "x2": 1, "x3": 0, "x4": 3}),
        dict(l=0, e=1, resB=[1, 0], resC=[3, -1], resV=2,
             resA=[["0.5", "0.5"], [2, -2]],
             resVariables={0: "x1", 1: "x3", 2: "x2", 3: "x4"},
             resCoefficients={"x1": 0, "x2": 2, "x3": 1, "x4": 3}),
        dict(l=1, e=0, resB=[1, 1], resC=[-1, -2], resV=4,
             resA=[["-0.25", 1], ["0.25", 1]],
             resVariables={0: "x4", 1: "x2", 2: "x3", 3: "x1"},
             resCoefficients={"x1": 3, "x2": 1, "x3": 2, "x4": 0}),
        dict(l=1, e=1, resB=[0, 1], resC=[2, "-0.5"], resV=2,
             resA=[[-1, "-0.5"], [1, "0.25"]],
             resVariables={0: "x1", 1: "x4", 2: "x3", 3: "x2"},
             resCoefficients={"x1": 0, "x2": 3, "x3": 2, "x4": 1}),
    ],
}

# LPStateSpec.groovy:102-131 — pivotConcurrently on 4x5
PIVOT_4x5 = {
    "A": [[1, 2, 4, 2, 2], [5, 5, 2, 1, 1], [2, 2, 1, 1, 4], [4, 2, 4, 1, 2]],
    "b": [2, 1, 4, 2], "c": [2, 4, 1, 5, 1], "m": 4, "n": 5,
    "cases": [
        dict(e=0, l=0, resB=[2, -9, 0, -6], resC=[-2, 0, -7, 1, -3], resV=4,
             resA=[[1, 2, 4, 2, 2], [-5, -5, -18, -9, -9], [-2, -2, -7, -3, 0], [-4, -6, -12, -7, -6]],
             resVariables={0: "x6", 1: "x2", 2: "x3", 3: "x4", 4: "x5", 5: "x1", 6: "x7", 7: "x8", 8: "x9"},
             resCoefficients={"x1": 5, "x2": 1, "x3": 2, "x4": 3, "x5": 4, "x6": 0, "x7": 6, "x8": 7, "x9": 8}),
        dict(e=2, l=2, resB=[-14, -7, 4, -14], resC=[0, 2, -1, 4, -3], resV=4,
             resA=[[-7, -6, -4, -2, -14], [1, 1, -2, -1, -7], [2, 2, 1, 1, 4], [-4, -6, -4, -3, -14]],
             resVariables={0: "x1", 1: "x2", 2: "x8", 3: "x4", 4: "x5", 5: "x6", 6: "x7", 7: "x3", 8: "x9"},
             resCoefficients={"x1": 0, "x2": 1, "x3": 7, "x4": 3, "x5": 4, "x6": 5, "x7": 6, "x8": 2, "x9": 8}),
    ],
}

# LPStateSpec.groovy:134-163 — pivotConcurrently on 7x2
PIVOT_7x2 = {
    "A": [[2, 4], [7, 2], [5, 4], [1, 3], [4, 1], [6, 2], [1, 7]],
    "b": [3, 6, 5, 10, 2, 1, 4], "c": [4, 3], "m": 7, "n": 2,
    "cases": [
        dict(e=0, l=3, resB=[-17, -64, -45, 10, -38, -59, -6], resC=[-4, -9], resV=40,
             resA=[[-2, -2], [-7, -19], [-5, -11], [1, 3], [-4, -11], [-6, -16], [-1, 4]],
             resVariables={0: "x6", 1: "x2", 2: "x3", 3: "x4", 4: "x5", 5: "x1", 6: "x7", 7: "x8", 8: "x9"},
             resCoefficients={"x1": 5, "x2": 1, "x3": 2, "x4": 3, "x5": 4, "x6": 0, "x7": 6, "x8": 7, "x9": 8}),
        dict(e=1, l=4, resB=[-5, 2, -3, 4, 2, -3, -10], resC=[-8, -3], resV=6,
             resA=[[-14, -4], [-1, -2], [-11, -4], [-11, -3], [4, 1], [-2, -2], [-27, -7]],
             resVariables={0: "x1", 1: "x7", 2: "x3", 3: "x4", 4: "x5", 5: "x6", 6: "x2", 7: "x8", 8: "x9"},
             resCoefficients={"x1": 0, "x2": 6, "x3": 2, "x4": 3, "x5": 4, "x6": 5, "x7": 1, "x8": 7, "x9": 8}),
    ],
}

# LPSolverSpec.groovy:9-21 — minInB
MIN_IN_B = [([1], 0), ([1, 0, -1], 2), ([1, 1, 1, 2], 0), ([-1, -1000, -10, -1001], 3), ([], -1)]

# LPSolverSpec.groovy:24-35 — x0 naming: result must not collide
X0_NAMES = [{"x1": 0, "x2": 1, "x3": 2}, {"x0": 0, "x1": 1}, {"auxVar": 0, "x0": 1, "auxVar1": 2}]

# LPSolverSpec.groovy:37-57 — aux-LP construction
AUX_CONSTRUCTION = {
    "A": [[1, 2, 3, 4, 5], [5, 4, 3, 2, 1], [1, 2, 3, 4, 5], [5, 4, 3, 2, 1], [1, 2, 3, 4, 5]],
    "b": [1, 2, 3, 4, 5], "c": [1, 2, 3, 4, 5], "m": 5, "n": 5,
    "resA": [[1, 2, 3, 4, 5, -1], [5, 4, 3, 2, 1, -1], [1, 2, 3, 4, 5, -1], [5, 4, 3, 2, 1, -1],
             [1, 2, 3, 4, 5, -1]],
    "resC": [0, 0, 0, 0, 0, -1],
}

# End-to-end known answers.  `names`: how many named structural variables the spec passes
# (None = the no-names constructor).  `log`: the (entering, leaving) pairs EXACTLY as
# LPState.java:115-118 logs them in logs/lp_solver.log (names when the state has names —
# the leaving entry is `variables.get(leaving)`, i.e. the name at non-basic position
# `leaving` — raw indices otherwise); None where the log does not hold the run.
SOLVE = [
    # LPSolverSpec.groovy:76-87 "lp solving [1]" -> 8 ; logs/lp_solver.log:456-493
    dict(name="lp_solving_1", A=[[4, -1], [2, 1], [-5, 2]], b=[8, 10, 2], c=[1, 1], m=3, n=2,
         maximize=True, names=["x1", "x2", "x3"], verdict="optimal", value="8.000000",
         aux_log=None, log=[("x1", "x1"), ("x2", "x2"), ("x4", "x1")]),
    # LPSolverSpec.groovy:89-98 "minimization" -> -17 ; logs/lp_solver.log:494-530
    dict(name="minimization", A=[[1, -4], [1, -1], [1, 1]], b=[0, 3, 11], c=[-3, 1], m=3, n=2,
         maximize=False, names=None, verdict="optimal", value="-17.000000",
         aux_log=None, log=[(0, 0), (1, 1), (0, 2)]),
    # LPSolverSpec.groovy:100-111 "initial infeasible solution" -> 20 ; logs/lp_solver.log:531-599
    dict(name="initial_infeasible", A=[[1, 0], [-1, 0], [0, 1], [0, -1]], b=[10, -2, 10, -2],
         c=[1, 1], m=4, n=2, maximize=True, names=["x1", "x2"], verdict="optimal",
         value="20.000000", x0_index=1,
         aux_log=[("x0", "x2"), ("x1", "x3"), ("x2", "x2")], log=[("x6", "x3"), ("x4", "x5")]),
    # LPSolverSpec.groovy:151-163 "unbounded linear program" ; logs/lp_solver.log:636-652
    dict(name="unbounded", A=[[1, 0]], b=[1], c=[1, 1], m=1, n=2, maximize=True, names=None,
         verdict="unbounded", message="This linear program is unbounded",
         aux_log=None, log=[(0, 0)]),
    # LPSolverSpec.groovy:165-177 "unbounded with initial infeasible solution" ; log :653-709
    dict(name="unbounded_after_phase1", A=[[-1, 0], [0, 1], [0, -1]], b=[-2, 4, -2], c=[1, 1],
         m=3, n=2, maximize=True, names=None, verdict="unbounded",
         message="This linear program is unbounded", x0_index=1,
         aux_log=[("x0", "x1"), ("x1", "x3"), ("x2", "x5")], log=[("x5", "x3")]),
    # LPSolverSpec.groovy:180-192 "infeasible linear program" row 1 ; log :710-741
    dict(name="infeasible_1", A=[[1], [-1]], b=[0, -1], c=[1], m=2, n=1, maximize=True,
         names=None, verdict="infeasible", message="This linear program is infeasible",
         x0_index=3, aux_log=[("x0", "x0"), ("x1", "x1")], log=None),
    # LPSolverSpec.groovy:180-192 row 2 ; log :742-763
    dict(name="infeasible_2", A=[[1, 1]], b=[-1], c=[1, 1], m=1, n=2, maximize=True, names=None,
         verdict="infeasible", message="This linear program is infeasible",
         x0_index=3, aux_log=[("x0", "x1")], log=None),
]

# LPSolverSpec.groovy:113-124 "auxiliary lp solving": solveAuxLP(auxLP, 2, 1) -> v == 0 ;
# logs/lp_solver.log:600-634 (x0 ends at index 1)
AUX_SOLVE = {
    "A": [[1, 0, -1], [-1, 0, -1], [0, 1, -1], [0, -1, -1]], "b": [10, -2, 10, -2],
    "c": [0, 0, -1], "m": 4, "n": 3,
    "variables": {0: "x1", 1: "x2", 2: "x0", 3: "x3", 4: "x4", 5: "x5", 6: "x6"},
    "index_of_x0": 2, "min_in_b": 1, "resV": 0, "x0_index": 1,
    "log": [("x0", "x2"), ("x1", "x3"), ("x2", "x2")],
}

# LPSolverSpec.groovy:126-149 "restoring initial lp"
RESTORE = {
    "A": [[0, -2, 1], [-1, 1, 0], [1, -2, 0], [0, 1, -1]], "b": [8, 2, 8, 2], "c": [0, -1, 0],
    "m": 4, "n": 3,
    "variables": {0: "x6", 1: "x0", 2: "x4", 3: "x3", 4: "x2", 5: "x5", 6: "x1"},
    "init_c": [1, 1], "init_variables": {0: "x1", 1: "x2"}, "init_m": 4, "init_n": 2,
    "index_of_x0": 1,
    "resA": [[0, 1], [-1, 0], [1, 0], [0, -1]], "resB": [8, 2, 8, 2], "resC": [1, 1], "resV": 4,
    "resVariables": {0: "x6", 1: "x4", 2: "x3", 3: "x2", 4: "x5", 5: "x1"},
    "resCoefficients": {"x6": 0, "x4": 1, "x3": 2, "x2": 3, "x5": 4, "x1": 5},
}

# io_files/input.txt:1-16 (LP #1, the only one readLP(File) reaches) -> io_files/output.txt:214-233
INPUT_TXT_LP1 = {
    "value": "7.000000",
    "primal": [1, 0, 0, 1, 0, 0, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 1],
}
