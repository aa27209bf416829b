"""profiles/traffic.json from the committed-round ncu --set full captures (gpurun_out/prof_head.ncu-rep, prof_eval.ncu-rep):
DRAM bytes (read + write) per launch of every kernel bench.py may report a roofline for."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = [('pool_tma_kernel', 'pool'), ('pool_kernel', 'pool'), ('graph_kernel', 'graph'), ('EpiGraphLayer', 'gemm_graph_layer'),
         ('attn_kernel', 'attn'), ('EpiDistance', 'gemm_distance'), ('rank_market_kernel', 'rank_market'), ('rank_mars_kernel', 'rank_mars')]
UNITS = {'pool': (882, 'tracklet'), 'graph': (882, 'tracklet'), 'gemm_graph_layer': (882, 'tracklet'), 'attn': (882, 'tracklet'),
         'gemm_distance': (1, '1980x9330x4096 matrix'), 'rank_market': (1, '1980x9330 matrix'), 'rank_mars': (1, '1980x9330 matrix')}


def rows(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    h, u = r[0], r[1]
    for x in r[2:]:
        yield {k: (v, uu) for k, uu, v in zip(h, u, x)}


def to_bytes(v, unit):
    return float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


res = {'_source': 'ncu --set full captures of round-1 v7 (profiles/r1/ncu_head_v7.txt, ncu_eval_v7.txt); bench.py --pool 882 / tools/run_eval_once.py'}
for rep in sys.argv[1:]:
    for row in rows(rep):
        name = row['Kernel Name'][0]
        key = next((k for pat, k in NAMES if pat in name), None)
        if key is None or key in res:
            continue
        b = to_bytes(*row['dram__bytes_read.sum']) + to_bytes(*row['dram__bytes_write.sum'])
        res[key] = {'dram_bytes_per_launch': b, 'units_per_launch': UNITS[key][0], 'unit': UNITS[key][1], 'kernel': name.strip()[:80]}
json.dump(res, open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w'), indent=1)
print(json.dumps(res, indent=1))
