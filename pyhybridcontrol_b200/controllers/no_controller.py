"""reference: controllers/no_controller.py:4"""
from .controller_base import ConstraintSolvedController


class NoController(ConstraintSolvedController):
    pass
