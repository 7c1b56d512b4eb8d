import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(f"{tag:28s} total {d['value']:6.2f}  fwd {d['fwd']['value']:6.2f} ({d['fwd']['ms']:7.2f} ms)  bwd {d['bwd']['value'] if d['bwd']['value'] and d['bwd']['ms']>0.01 else 0:6.2f} ({d['bwd']['ms']:7.2f} ms)  brick {d['phase_ms_per_step']['brick']:.2f} ms")
except Exception as e:
    print(tag, "FAILED", e); print(open(path).read()[-800:])
