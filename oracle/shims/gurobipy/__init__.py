"""Shim (oracle only) for `import gurobipy as gp; from gurobipy import GRB` (utils.py:41-42).

The two Gurobi QPs (utils.py:592-650) are replaced in ref_oracle.py by exact vertex enumeration of the
input polytope (the maximum of a convex quadratic over a polytope is attained at a vertex).
"""


class GRB:
    INFINITY = float("inf")
    MAXIMIZE = -1
    MINIMIZE = 1


def Model(*a, **k):
    raise RuntimeError("gurobipy shim: bar_u_solve/bar_d_u_solve must be patched (see oracle/ref_oracle.py)")


def quicksum(it):
    return sum(it)
