"""Install the UNMODIFIED reference next to the repo so that it travels to the GPU box: baseline/_ref/ (git-ignored).

The reference is four Python files plus a prebuilt NNs.so and data files; it has no setup.py, so the "offline install"
of the bench contract is a plain copy of the files its selection path needs:
    cut_select_qp.py, cut_select_qcqp.py, utilities.py, neural_nets/NNs.so,
    boxqp_instances/{spar020-100-1, spar030-060-1, spar125-075-1}.in, qcqp_instances/q_20_20_100_1.osil
Nothing under baseline/_ref is imported by the product; bench.py's `cpu_baseline_reference` leg and the drop-in tests
(tests/test_gpu_reference_loop.py) import it with stub cplex / mosek / cvxopt / chompack / lxml modules.

    python tools/install_reference.py [/root/reference]
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["cut_select_qp.py", "cut_select_qcqp.py", "utilities.py", "LICENSE", "neural_nets/NNs.so",
         "boxqp_instances/spar020-100-1.in", "boxqp_instances/spar030-060-1.in", "boxqp_instances/spar125-075-1.in",
         "qcqp_instances/q_20_20_100_1.osil"]


def install(src="/root/reference"):
    if not os.path.isdir(src):
        return False
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DEST, rel)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s):
            shutil.copy2(s, d)
    return True


if __name__ == "__main__":
    ok = install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("installed into %s" % DEST if ok else "reference tree not found")
