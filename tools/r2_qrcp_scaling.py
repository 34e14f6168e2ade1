"""device QRCP of a 1024 x 44032 matrix (the SE-QRCS candidate step of config E): time against the number of elimination steps"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import itcpd
m, n = 1024, int(os.environ.get("QR_N", 44032))
steps_list = [int(x) for x in os.environ.get("QR_STEPS", "1,128,256,512,768,1024").split(",")]
rng = np.random.default_rng(0)
A = np.asfortranarray(rng.standard_normal((m, n)))
eng = itcpd.Engine(0)
eng.qrcp_matrix(A, steps=1)
base = None
for s in steps_list:
    best = 1e9
    for rep in range(2 if len(steps_list) > 1 else 1):
        eng.synchronize(); t0 = time.perf_counter(); piv, rd = eng.qrcp_matrix(A, steps=s); eng.synchronize()
        best = min(best, time.perf_counter() - t0)
    if base is None:
        base = best
    print(f"steps {s:5d}: {1e3 * best:8.2f} ms   minus upload {1e3 * (best - base):8.2f} ms", flush=True)
if os.environ.get("QR_CHECK"):
    import scipy.linalg
    R, p = scipy.linalg.qr(A[:, :4096], mode="r", pivoting=True)
    piv, rd = eng.qrcp_matrix(np.asfortranarray(A[:, :4096]))
    print("pivots equal lapack:", int(np.sum(piv[:1024] - 1 == p[:1024])), "of 1024; rdiag rel", float(np.max(np.abs(np.abs(rd) - np.abs(np.diag(R))) / np.abs(R[0, 0]))))
eng.close()
