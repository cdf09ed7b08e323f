"""Throughput of the diagonal k sampler (for timing and ncu):

    python tests/tools/prof_diagk.py [m sigma l n]       # JSON line
    ncu --set full --clock-control none --import-source on -k regex:k_diagk$ -c 1 \
        -o gpurun_out/diagk python tests/tools/prof_diagk.py

Device-resident timing with CUDA events on the launching stream (qb200_diagk_sample_device), the
end-to-end time of the synchronous C ABI with host rows (qb200_diagk_sample), and a parity check
of a sub-sample against the CPU twin (tests/hostsim) -- plus, with --ref, the reference's own
sample_k_from_diagonal_j_eta_pivot timed on one host core (oracle/_ref).
"""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import qunundrum_b200 as qb  # noqa: E402

LD = np.longdouble


def batch(m, sigma, n, seed):
    prng = random.Random(seed)
    r = (1 << (m - 1)) + 1 + prng.randrange((1 << (m - 1)) - 1)
    d = r // 2 + prng.randrange(r // 2)
    wj = (m + sigma + 31) // 32
    rng = np.random.default_rng(seed)
    J = rng.integers(0, 1 << 32, size=(n, wj), dtype=np.uint64).astype(np.uint32)
    if (m + sigma) % 32:
        J[:, -1] &= np.uint32((1 << ((m + sigma) % 32)) - 1)
    eta = rng.integers(-25, 26, size=n).astype(np.int32)
    piv = rng.random(n).astype(LD)
    return d, r, J, eta, piv


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    m, sigma, l, n = (int(a) for a in args) if len(args) == 4 else (2048, 5, 2048, 296 * 1024)
    delta_bound = 1000
    d, r, J, eta, piv = batch(m, sigma, n, 11)
    ctx = qb.Context(0)
    S = qb.DiagonalKSampler(qb.Diagonal_Parameters(m, sigma, 0, d, r, eta_bound=25, t=30, l=l), ctx)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    dJ = torch.from_numpy(J.view(np.int32)).cuda()
    dE = torch.from_numpy(eta).cuda()
    dP = torch.from_numpy(piv.view(np.uint8).reshape(n, 16)).cuda()
    dK = torch.zeros((n, S.k_limbs), dtype=torch.int32, device="cuda")
    dO = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
    times = []
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        S.sample_device(n, dJ.data_ptr(), dE.data_ptr(), dP.data_ptr(), delta_bound, dK.data_ptr(),
                        dO.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        stream.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times[2:]))
    out = dO.cpu().numpy()
    status = out[:, 3].view(np.int64) & 0xffffffff
    delta = out[:, 2].view(np.int64)
    res = {"workload": f"sample_k_from_diagonal_j_eta_pivot, m={m} sigma={sigma} l={l}, {n} samples, "
                       f"delta_bound={delta_bound}, uniform j and pivots",
           "samples": n, "ms": ms, "samples_per_s": n / ms * 1e3, "ms_all": times,
           "ok_fraction": float((status == 0).mean()), "mean_abs_delta": float(np.abs(delta).mean()),
           "max_abs_delta": int(np.abs(delta).max())}
    # integer work: multiply-adds of the five products per sample (diagk.cuh)
    k = (m + 31) // 32
    wj = S.j_limbs
    mads = k * wj + k * k + 2 * ((k + 1) * (k + 2) - k * (k - 1) // 2 + k * (k + 1) // 2 + k)
    res["imad_per_sample"] = mads
    res["imad_per_s"] = mads * res["samples_per_s"]
    t0 = time.perf_counter()
    ks, x, dl, st = S.sample(J, eta, piv, delta_bound)
    t1 = time.perf_counter()
    res["e2e"] = {"ms": (t1 - t0) * 1e3, "samples_per_s": n / (t1 - t0),
                  "note": "qb200_diagk_sample with host rows, Python int conversion of k included"}
    t0 = time.perf_counter()
    S.sample(J, eta, piv, delta_bound, want_k=False)
    t1 = time.perf_counter()
    res["e2e_no_k"] = {"ms": (t1 - t0) * 1e3, "samples_per_s": n / (t1 - t0)}
    # parity of a sub-sample against the CPU twin
    from tests import hostsim as hs
    sub = np.random.default_rng(3).choice(n, 400, replace=False)
    T = hs.DiagK(m, sigma, l, d, r)
    ks2, x2, dl2, st2 = T.sample([hs.limbs_to_int(J[i]) for i in sub], eta[sub], piv[sub], delta_bound)
    res["twin_parity"] = bool([ks[i] for i in sub] == ks2 and np.array_equal(dl[sub], dl2)
                              and np.array_equal(st[sub], st2))
    gk = dK.cpu().numpy().view(np.uint32)
    res["device_form_equals_host_form"] = bool(all(hs.limbs_to_int(gk[i]) == ks[i] for i in sub))
    if "--ref" in sys.argv:
        from oracle import ref as R
        P = R.RefDiagonalParameters(m, sigma, 0, d, r, eta_bound=25, t=30, l=l)
        cnt, t0 = 0, time.perf_counter()
        same = True
        for i in sub[:200]:
            ok, kk, _, _ = R.sample_k_from_diagonal_j_eta_pivot(P, piv[i], hs.limbs_to_int(J[i]), int(eta[i]),
                                                                delta_bound, precision=256)
            same = same and kk == ks[i] and ok == (st[i] == 0)
            cnt += 1
            if time.perf_counter() - t0 > 20:
                break
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": cnt / dt, "unit": "samples/s", "cores": 1, "kind": "reference",
                               "sample": f"{cnt} calls of the reference's sample_k_from_diagonal_j_eta_pivot, {dt:.1f} s",
                               "identical_k": bool(same)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
