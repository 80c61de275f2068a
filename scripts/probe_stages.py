"""Development probe: per-GPU shard of an 8-GPU run on one GPU: stream-kernel time against ring depth / options.
usage: CONFIGS="stages=8,stages=10+merge_depth=4" python scripts/probe_stages.py [nx ndom world rank]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import krylov_b200  # noqa: F401
from krylov_b200 import fem
from krylov_b200.hostprep import problem as PB

nx, ndom, w, r = (int(v) for v in (sys.argv[1:5] + ["2001", "256", "8", "3"][len(sys.argv) - 1:]))
base = PB.build_mesh_and_partition(nx, ndom)
sd = base[-1]
d0, d1 = (sd.ndom * r) // w, (sd.ndom * (r + 1)) // w
pr = PB.realization(base, ireal=1, dom_range=(d0, d1))
for cfg in os.environ.get("CONFIGS", "stages=6").split(","):
    opts = {k: int(v) for k, v in (kv.split("=") for kv in cfg.split("+") if kv)}
    o = {"shard_world": w, "shard_rank": r}
    o.update(opts)
    try:
        S = fem.B200SchurOperator(sd, pr.blocks, pr.factors, options=o)
        S.h.finalize()
        st = S.h.stats()
        line = (f"{cfg:40s} warps {st['consumer_warps']:.0f} stages {st['ring_stages']:.0f} md {st['merge_depth']:.0f} "
                f"levels {st['max_levels']:.0f} smem {st['smem_bytes'] / 1024:.0f} KB stream {st['stream_bytes'] / 1e9:.3f} GB:")
        line += f"  read {st['stream_read_bytes'] / 1e9:.3f} GB  auto pf {st['l2_prefetch_chunks']:.0f}: {S.h.time_kernel(0, 50) * 1e3:.1f} us"
        for mode in (0, 1):
            S.h.set_option("l2_prefetch_mode", mode)
            for pf in [int(v) for v in os.environ.get("PF", "").split(",") if v]:
                S.h.set_option("l2_prefetch", pf)
                line += f" | m{mode} pf={pf} {S.h.time_kernel(0, 50) * 1e3:.1f}"
        S.h.set_option("l2_prefetch_mode", 0)
        S.h.set_option("l2_prefetch", int(st['l2_prefetch_chunks']))
        print(line, flush=True)
        if os.environ.get("DBG"):
            S.h.set_option("debug_timing", 1)
            S.h.time_kernel(0, 1)
            cs_ = int(st["csize"])
            slots = int(os.environ.get("KB200_DBG_SLOTS", "64"))
            st_ = S.h.debug_stamps(int(st["n_cta"]), slots).astype(float)
            names = [("init", 0, 9), ("FWD", 9, 10), ("PART", 10, 1), ("wait sync1", 1, 2), ("top rhs + TOPF", 2, 3), ("TOPB", 3, 4),
                     ("BWD", 5, 6), ("GG+OUT", 6, 7), ("epilogue", 7, 8), ("total", 0, 8)]
            t0 = st_[:, 0].min()
            print(f"   CTA start spread {(st_[:, 0].max() - t0) / 1.9e3:.1f} us, end spread {(st_[:, 8].max() - st_[:, 8].min()) / 1.9e3:.1f} us, "
                  f"first start -> last end {(st_[:, 8].max() - t0) / 1.9e3:.1f} us")
            for nm, i, j in names:
                a = (st_[:, j] - st_[:, i]) / 1.9e3
                r0 = a[0::cs_]
                print(f"   {nm:16s} all CTAs mean {a.mean():7.1f} min {a.min():7.1f} max {a.max():7.1f} us | rank-0 CTAs mean {r0.mean():7.1f} us")
            if slots > 64:   # instrumented build: per-level clocks of consumer warp 0 (negative = a section starts)
                for cta in (0, 1, cs_ * (int(st["n_cta"]) // cs_ // 2)):
                    lv = st_[cta, 64:]
                    lv = lv[lv != 0]
                    out, sec, prev = [], -1, None
                    for v in lv:
                        if v < 0:
                            sec += 1
                            prev = -v
                            out.append(f"| S{sec}:")
                        else:
                            out.append(f"{(v - prev) / 1.9e3:.2f}")
                            prev = v
                    print(f"   CTA {cta}: level times (us) per section: {' '.join(out)}")
                pf = st_[:, 16:48].reshape(-1, 8, 4)
                tot = (st_[:, 7] - st_[:, 0]).mean() / 1.9e3
                print(f"   warps 0-7, mean over CTAs: slices {pf[:, :, 1].mean():.0f}, in slices {pf[:, :, 0].mean() / 1.9e3:.1f} us "
                      f"(slowest warp {pf[:, :, 0].max(axis=1).mean() / 1.9e3:.1f}), cycles/slice {pf[:, :, 0].sum() / max(1, pf[:, :, 1].sum()):.0f}, "
                      f"waiting for chunks {pf[:, :, 3].mean() / 1.9e3:.1f} us, kernel {tot:.1f} us")
            S.h.set_option("debug_timing", 0)
        del S
    except Exception as e:
        print(f"{cfg:40s} failed: {e}", flush=True)
