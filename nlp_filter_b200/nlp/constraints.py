"""Constraint functions under the reference's names (nlp/constraints.py:4-5)."""


def equality_constaint(x, params=None):
    """x[0] - x[1] (== 0); spelling as in the reference."""
    return x[0] - x[1]
