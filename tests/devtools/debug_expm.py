import sys, os, ctypes
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
from chromevol_enhanced_b200 import synthetic, _lib
from chromevol_enhanced_b200.engine import ExpmWorkspace, _stream_ptr
from oracle import chromevol_oracle as O
from oracle.expm_amh import expm_amh
from device_model import prepare, select, expm_model, postprocess
np.set_printoptions(linewidth=200, precision=4)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 35
n = N + 1
rows = synthetic.random_param_rows(24, seed=4, linear='positive'); rows[::3, 8] = 7
times = np.array([0.0537514513, 0.2, 0.01])
lib = _lib.load(); dev = torch.device("cuda:0")
B, E = rows.shape[0], len(times)
ws = ExpmWorkspace(n, E, B, dev)
ws.params.copy_(torch.as_tensor(rows, device=dev)); d_t = torch.as_tensor(times, device=dev)
st = _stream_ptr()
lib.cevb_build_q(_lib.ptr(ws.params), B, n, _lib.ptr(ws.pw), st)
lib.cevb_expm_prepare(_lib.ptr(ws.pw), B, n, _lib.ptr(ws.norms), _lib.ptr(ws.csc), _lib.ptr(ws.csc_cnt), st)
for fp in (0, 1):
    lib.cevb_expm_branches(_lib.ptr(ws.pw), _lib.ptr(ws.norms), _lib.ptr(ws.csc), _lib.ptr(ws.csc_cnt), _lib.ptr(d_t), B, E, n,
                           _lib.ptr(ws.P), _lib.ptr(ws.info), _lib.ptr(ws.scratch), _lib.ptr(ws.counter), fp, st)
    torch.cuda.synchronize()
    pw = ws.pw.cpu().numpy()[:, :, :, :n]; norms = ws.norms.cpu().numpy(); P = ws.P.cpu().numpy()[:, :, :, :n]; info = ws.info.cpu().numpy()
    cnt = ws.csc_cnt.cpu().numpy(); csc = ws.csc.cpu().numpy().astype(np.int64) & 0xffff
    for b in range(0, 6):
        Q = O.build_q(synthetic.row_to_dict(rows[b]), N)
        powers, nm = prepare(Q)
        perr = [np.abs(pw[b, k] - powers[2 * k]).max() / max(1e-300, np.abs(powers[2 * k]).max()) for k in range(1, 7)]
        nref = [nm['n1'], nm['d4'], nm['d6'], nm['d8'], nm['d10'], nm['N7'], nm['N11'], nm['N15'], nm['N19'], nm['N27']]
        nerr = np.abs(norms[b, :10] - nref) / np.maximum(np.abs(nref), 1e-300)
        # csc check
        bad = 0
        for j in range(n):
            nzr = np.flatnonzero(Q[:, j])
            if cnt[b, j] != len(nzr) or not np.array_equal(csc[b, :len(nzr), j], nzr): bad += 1
        print(f"fp={fp} b={b} Qerr={np.abs(pw[b,0]-Q).max():.1e} powers_relerr={['%.1e'%x for x in perr]} norms_relerr_max={nerr.max():.1e} csc_bad={bad} dense={norms[b,10]}")
        for e, t in enumerate(times):
            ref = O.transition_probabilities(Q, t)
            Xm, mm, sm = expm_model(Q, t, (powers, nm)); 
            _, m2, s2 = expm_amh(Q * t, True)
            err = np.abs(P[b, e] - ref).max(); errm = np.abs(postprocess(Xm) - ref).max()
            print(f"    t={t:.4f} dev(m,s)=({info[b,e]&255},{(info[b,e]>>8)&255},fl={info[b,e]>>16}) model=({mm},{sm}) amh=({m2},{s2}) err_dev={err:.2e} err_model={errm:.2e}")
            if err > 1e-6 and b == 0 and e == 0:
                print("dev row5", P[b, e, 5, :12]); print("ref row5", ref[5, :12])
                print("dev row20", P[b, e, 20, 14:28]); print("ref row20", ref[20, 14:28])
