#!/usr/bin/env python
"""Compare the reference-format debug files (send<r>.txt / recv<r>.txt, event.cpp:203-227 formats) of two runs,
e.g. the p2p backend on GPUs against the gloo backend on CPUs: fired / new-value flags must agree, norms and
thresholds to a relative tolerance (different conv arithmetic => last-digit differences in the printed %g values)."""
import glob, json, os, sys


def parse(path):
    rows = []
    for ln in open(path):
        f = [x.strip() for x in ln.strip().split(",") if x.strip() != ""]
        rows.append([float(x) for x in f])
    return rows


def main(a, b, tol=2e-4):
    out = {"files": 0, "rows": 0, "values": 0, "flag_mismatch": 0, "value_mismatch": 0, "max_rel": 0.0, "missing": []}
    for fa in sorted(glob.glob(os.path.join(a, "*.txt"))):
        name = os.path.basename(fa)
        if not (name.startswith("send") or name.startswith("recv")):
            continue
        fb = os.path.join(b, name)
        if not os.path.exists(fb):
            out["missing"].append(name)
            continue
        ra, rb = parse(fa), parse(fb)
        out["files"] += 1
        if len(ra) != len(rb):
            out["missing"].append(f"{name}: {len(ra)} vs {len(rb)} rows")
        for x, y in zip(ra, rb):
            out["rows"] += 1
            for u, v in zip(x, y):
                out["values"] += 1
                if u in (0.0, 1.0) and v in (0.0, 1.0):
                    out["flag_mismatch"] += int(u != v)
                else:
                    rel = abs(u - v) / max(abs(u), abs(v), 1e-12)
                    out["max_rel"] = max(out["max_rel"], rel)
                    out["value_mismatch"] += int(rel > tol)
    print(json.dumps(out))
    return out


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 2e-4)
