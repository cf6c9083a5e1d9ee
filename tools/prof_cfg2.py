import cProfile, pstats, io, os, sys, time
import numpy as np
ROOT="/root/repo"
sys.path.insert(0, ROOT)
import sdpcutsel_via_nn_b200 as pkg
g = np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))
Qf = g["inst_spar125_075_1_Q"].astype(np.float64)
Q_arr, adj = pkg.synthetic.boxqp_arrays(Qf)
n = Qf.shape[0]; vv = g["cfg2_vars"]
cs = pkg.CutSolver(); cs.set_instance(Q_arr, adj, n, dim=3); cs._load_neural_nets()
N = cs._get_sdp_vertex_cover(3); k = min(int(np.floor(0.1 * N)), 5000)
cs._CutSolver__preprocess_triangle_ineq()
def rnd(strat):
    cs._my_prob.linear_constraints.rows = []
    r = cs._sel_eigcut_by_ordering_on_measure(strat, vv, 1, sel_size=k)
    rl = r[1] if strat == 4 else r
    cs._gen_eigcuts_selected(strat, k, rl, vars_values=vv)
    cs._CutSolver__separate_and_add_triangle(0.1, vv)
for s in (1,2,4): rnd(s); rnd(s)
import gc; gc.disable()
for s in (4,):
    t=time.perf_counter(); 
    for _ in range(10): rnd(s)
    print("strat",s,"round ms",(time.perf_counter()-t)*100)
    pr=cProfile.Profile(); pr.enable()
    for _ in range(10): rnd(s)
    pr.disable()
    st=io.StringIO(); pstats.Stats(pr,stream=st).sort_stats("cumulative").print_stats(35); print(st.getvalue()[:6000])
