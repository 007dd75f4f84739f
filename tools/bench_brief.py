"""One line per bench JSON file: step, stage times, e2e (for reading gpurun_out/ quickly)."""
import json, sys
for f in sys.argv[1:]:
    try:
        for line in open(f):
            line = line.strip()
            if not line.startswith('{'): continue
            d = json.loads(line)
            if 'metric' not in d: continue
            c = d.get('config', {}); s = d.get('stage_ms', {}) or {}; e = d.get('e2e', {}) or {}
            print(f"{f.split('/')[-1]:34s} N={d.get('n_gpus')} cfg={c.get('named_config')} ms={d.get('ms_per_step', 0):.2f} val={d.get('value', 0) / 1e6:.1f}M "
                  f"k1={s.get('k1', 0):.2f} k2e={s.get('k2e', 0):.2f} k2d={s.get('k2d', 0):.2f} k3={s.get('k3', 0):.2f} enc={s.get('enc_total', 0):.2f} dec={s.get('dec_total', 0):.2f} "
                  f"e2e={e.get('ms_per_step', 0) or 0:.2f} ({e.get('compress_ms', 0) or 0:.2f}+{e.get('decompress_ms', 0) or 0:.2f}) bpb={d.get('bits_per_base', 0):.5f} launches={d.get('gpu_launches')}")
    except Exception as ex:
        print(f, 'ERR', ex)
