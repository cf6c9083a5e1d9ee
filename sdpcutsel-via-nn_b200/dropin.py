"""Composition of the GPU selection path with the reference's own classes.

``make_solvers(cut_select_qp, cut_select_qcqp)`` returns subclasses of the reference's ``CutSolver`` / ``CutSolverQCQP``
whose hot-path methods are the B200 ones while ``cut_select_algo`` (the CPLEX LP loop), the instance readers and the
McCormick rows stay the reference's.  Three things a hand-written ``class CutSolver(B200CutSelection, ref.CutSolver)``
gets wrong are taken care of here:

* the reference's private methods are name-mangled with ITS class names (``_CutSolver__separate_and_add_triangle``,
  ``_CutSolver__gen_dense_eigcuts``, ``_CutSolverQCQP__get_vertex_cover``): the overrides must carry exactly those names;
* ``cut_select_algo`` reads ``CutSolver._THRES_MAX_SUBS`` from the reference's module-level class (cut_select_qp.py:117),
  not from ``self``: the 4e6 wall has to be lifted THERE for covers the reference could not hold in RAM;
* the QCQP class must come first in the MRO (its ``cut_select_algo`` calls ``super()._sel_eigcut...``), so its O(N^2)
  ``__get_vertex_cover`` has to be overridden explicitly with the key-based cover algebra.
"""
from .cut_select_qp import B200CutSelection
from .cut_select_qcqp import two_pattern_covers


def make_solvers(ref_qp, ref_qcqp=None, lift_subproblem_wall=True):
    """ref_qp / ref_qcqp: the imported reference modules cut_select_qp / cut_select_qcqp.
    Returns (CutSolver, CutSolverQCQP or None)."""

    class CutSolver(B200CutSelection, ref_qp.CutSolver):
        def _CutSolver__preprocess_triangle_ineq(self):
            return self._tri_preprocess()

        def _CutSolver__separate_and_add_triangle(self, sel_size, vars_values):
            return self._tri_separate(sel_size, vars_values)

        def _CutSolver__gen_dense_eigcuts(self, vars_values=None):      # strat 0
            return self._dense_eigcuts(vars_values)

    if lift_subproblem_wall:
        # nothing is materialised per sub-problem on the GPU path; the guard at cut_select_qp.py:117 reads the class attribute
        ref_qp.CutSolver._THRES_MAX_SUBS = B200CutSelection._THRES_MAX_SUBS
    if ref_qcqp is None:
        return CutSolver, None

    class _Mix(B200CutSelection, ref_qp.CutSolver):
        pass

    # MRO: CutSolverQCQP -> reference CutSolverQCQP -> _Mix -> B200CutSelection -> reference CutSolver: the super() calls in
    # cut_select_qcqp.py:39-41, 66-77 land on the GPU path
    class CutSolverQCQP(ref_qcqp.CutSolverQCQP, _Mix):
        def _CutSolverQCQP__get_vertex_cover(self, dim):                # cut_select_qcqp.py:314-334
            return two_pattern_covers(self, dim)

    return CutSolver, CutSolverQCQP
