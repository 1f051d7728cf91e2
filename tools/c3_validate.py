"""At-scale validation of BASELINE.json configs[2]-like inputs on one GPU.
  python tools/c3_validate.py <n_bases> <n_records> [invert]
Builds the BWT, prints per-phase stats and the SHA-256 of the packed words; with `invert`, rebuilds the
text from the BWT by LF-walk (oracle) and compares it with the input (size-independent property)."""
import hashlib, json, os, sys, time
sys.path.insert(0, ".")
import numpy as np
from debwt_b200 import api, synth

n, nrec = int(float(sys.argv[1])), int(sys.argv[2])
invert = len(sys.argv) > 3 and sys.argv[3] == "invert"
t0 = time.time(); recs = synth.config4(n // nrec, nrec) if os.environ.get('DEBWT_CFG') == 'c4' else synth.config3(n, nrec); tg = time.time() - t0
print(f"generated {sum(r.size for r in recs)} bases in {len(recs)} records in {tg:.1f} s", flush=True)
with api.BwtBuilder() as b:
    for rep in range(int(os.environ.get("DEBWT_REPS", "2"))):
        t0 = time.time(); b.set_records(recs); b.build(); w, s, d = b.result(); wall = time.time() - t0
    st = b.stats()
st["wall_s"] = wall
st["Mbp_s_wall"] = n / wall / 1e6
st["Mbp_s_device"] = n / st["ms_total"] / 1e3
st["sort_GBps_136"] = 136 * st["n_keys"] / st["ms_sort"] / 1e6
st["sha256"] = hashlib.sha256(w.tobytes()).hexdigest()
st["sharp_sha256"] = hashlib.sha256(s.tobytes()).hexdigest()
st["dollar"] = int(d[0])
print(json.dumps(st), flush=True)
if invert:
    from oracle import coracle, stages as st_
    t0 = time.time()
    N = st["n_symbols"]
    sym = coracle.unpack_bwt(w, N, s, d)
    ok, text = coracle.invert_bwt(sym)
    del sym
    # expected text symbols
    exp = np.empty(N, dtype=np.uint8); pos = 0
    for r in recs:
        exp[pos:pos + r.size] = np.searchsorted(np.frombuffer(b"ACGT", dtype=np.uint8), r); pos += r.size
        exp[pos] = 4; pos += 1
    exp[-1] = 5
    same = bool(ok and (text == exp).all())
    print(json.dumps({"lf_inversion_ok": same, "invert_s": time.time() - t0}), flush=True)
    sys.exit(0 if same else 1)
