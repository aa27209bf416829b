"""profiles/traffic.json from ncu --set full captures (read here, no GPU needed): DRAM bytes (dram__bytes_read.sum +
dram__bytes_write.sum) of every head kernel summed over ONE 882-tracklet head call (both layers' launches of a name
together), and per launch for the evaluation kernels.   python tools/make_traffic.py HEAD.ncu-rep [EVAL.ncu-rep]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEAD = [('pool_tma_kernel', 'pool'), ('pool_kernel', 'pool'), ('graph_kernel_tc', 'graph'), ('EpiGraphLayer', 'gemm_graph_layer'),
        ('EpiPlain', 'gemm_graph_layer'), ('graph_mix_kernel', 'graph_mix'), ('attn_kernel', 'attn')]
EVAL = [('EpiDistance', 'gemm_distance', '1980x9330x4096 matrix'), ('rank_market_kernel', 'rank_market', '1980x9330 matrix'),
        ('rank_mars_kernel', 'rank_mars', '1980x9330 matrix'), ('split_planes', 'split_planes', '11310x4096 rows')]


def rows(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    h, u = r[0], r[1]
    for x in r[2:]:
        yield {k: (v, uu) for k, uu, v in zip(h, u, x)}


def to_bytes(v, unit):
    return float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


res = {'_source': 'ncu --set full --clock-control none, one 882-tracklet head call (tools/ncu_head_r2.sh) and tools/run_eval_once.py; '
                  'summaries in profiles/r2/ncu_head_final.txt, ncu_eval_final.txt'}
for row in rows(sys.argv[1]):
    name = row['Kernel Name'][0]
    key = next((k for pat, k in HEAD if pat in name), None)
    if key is None:
        continue
    b = to_bytes(*row['dram__bytes_read.sum']) + to_bytes(*row['dram__bytes_write.sum'])
    e = res.setdefault(key, {'dram_bytes_per_call': 0.0, 'units_per_call': 882, 'unit': 'tracklet', 'launches_per_call': 0, 'kernels': []})
    e['dram_bytes_per_call'] += b
    e['launches_per_call'] += 1
    e['kernels'].append(name.strip()[:70])
if len(sys.argv) > 2:
    for row in rows(sys.argv[2]):
        name = row['Kernel Name'][0]
        hit = next(((k, u) for pat, k, u in EVAL if pat in name), None)
        if hit is None or hit[0] in res:
            continue
        b = to_bytes(*row['dram__bytes_read.sum']) + to_bytes(*row['dram__bytes_write.sum'])
        res[hit[0]] = {'dram_bytes_per_call': b, 'units_per_call': 1, 'unit': hit[1], 'launches_per_call': 1, 'kernels': [name.strip()[:70]]}
json.dump(res, open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w'), indent=1)
print(json.dumps(res, indent=1))
