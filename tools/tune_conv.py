"""Per-layer sweep of drp_conv3x3's tile rows / shared-memory budget (env DRP_CONV_ROWS, DRP_CONV_SMEM_KB) at the U-Net's layer shapes."""
import itertools, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffrp_b200 import denoiser as dn, _abi

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
res = (1, 1, 2, 4, 8, 16, 16, 8, 8, 4, 4, 2, 2, 1, 1, 1)
modes = dict(enc_conv1=1, enc_conv2=1, enc_conv3=1, enc_conv4=1, enc_conv5b=2, dec_conv4b=2, dec_conv3b=2, dec_conv2b=2)
out = {}
for (name, cin, cout), r in zip(dn.LAYERS, res):
    H = W = R // r
    cb = dn._pad16(cin)
    x = torch.randn(H, W, cb, device='cuda')
    w = torch.randn(cout, cin, 3, 3, device='cuda') * 0.05
    b = torch.zeros(cout, device='cuda')
    wm, bm = dn.pack_weight(w, b, list(range(cin)), cb)
    mode = modes.get(name, 0)
    oh, ow = {0: (H, W), 1: (H // 2, W // 2), 2: (2 * H, 2 * W)}[mode]
    cs = (cout + 3) // 4 * 4
    y = torch.empty(oh, ow, dn._pad16(cout), device='cuda')
    best = {}
    combos = [(0, 0)] + list(itertools.product((8, 16), (24, 32, 48, 72, 100)))
    if os.environ.get('TUNE_PERSISTENT'):
        combos = [(0, 0), (1, 0), (1, 8), (1, 16)]            # (persistent?, rows) with the default budgets
    if os.environ.get('TUNE_KC'):
        combos = [(pp * 100 + kc, rows) for pp in (0, 1) for kc in (16, 32) for rows in (8, 16)]   # persistent x chunk x rows
    for rows, kb in combos:
        for k_ in ('DRP_CONV_ROWS', 'DRP_CONV_SMEM_KB', 'DRP_CONV_PERSISTENT'):
            os.environ.pop(k_, None)
        if os.environ.get('TUNE_KC'):
            os.environ['DRP_CONV_PERSISTENT'], os.environ['DRP_CONV_KC'], os.environ['DRP_CONV_ROWS'] = str(rows // 100), str(rows % 100), str(kb)
        elif os.environ.get('TUNE_PERSISTENT'):
            os.environ['DRP_CONV_PERSISTENT'] = str(rows)
            if kb:
                os.environ['DRP_CONV_ROWS'] = str(kb)
        elif rows:
            os.environ['DRP_CONV_ROWS'], os.environ['DRP_CONV_SMEM_KB'], os.environ['DRP_CONV_PERSISTENT'] = str(rows), str(kb), '0' 
        for _ in range(3):
            dn.conv3x3(x, 0, cb, wm, bm, y, 0, cs, mode, True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dn.conv3x3(x, 0, cb, wm, bm, y, 0, cs, mode, True)
        e1.record(); torch.cuda.synchronize()
        best[(rows, kb)] = e0.elapsed_time(e1) / 20 * 1e3
    k = min(best, key=best.get)
    out[name] = dict(best=k, us=round(best[k], 1), default=round(best.get((0, 0), best[k]), 1), all={"%d/%d" % kk: round(v, 1) for kk, v in best.items()})
    print(name, out[name]['best'], out[name]['us'], 'all', out[name]['all'], flush=True)
json.dump(out, open('gpurun_out/tune_conv.json', 'w'), indent=1)
print('sum best', sum(v['us'] for v in out.values()), 'sum default', sum(v['default'] for v in out.values()))
