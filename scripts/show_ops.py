import json, sys
for path in sys.argv[1:]:
    for l in open(path):
        if not l.startswith("{"):
            continue
        r = json.loads(l)
        if len(sys.argv) > 2 and "conv" not in r["op"]:
            continue
        print(f"{path.split('/')[-1][:18]:18s} {r['op'][:44]:44s} n={r['n']:8d} nv={r['nv']:8d} V={r.get('val_dim','-'):>4} {r['us']:10.1f}us GB/s={r.get('GBps',0):8.1f} hbm={r.get('hbm_frac',0):.3f} TF={r.get('TFLOPs',0):7.2f} ref_us={r.get('ref_us',0):10.1f}")
