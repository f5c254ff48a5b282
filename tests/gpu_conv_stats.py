"""Debug helper (GPU box): per-role cycle breakdown of the CTA-pair conv kernel on the dominant 280 -> 280 layer."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import _gpu_util as u
B, H, W, cin, cout = (int(a) for a in (sys.argv[1:6] if len(sys.argv) > 5 else (1, 512, 512, 280, 280)))
ctype = int(sys.argv[6]) if len(sys.argv) > 6 else 0
cin_pad, n_pad = u.pad16(cin), u.pad16(cout)
n_slots = B * (H + 1) * (W + 1)
x = (torch.randn((n_slots, cin_pad), device='cuda') * 0.5).to(torch.float16)
w = u.pack_weight((np.random.RandomState(0).normal(0, 0.03, (cout, cin, 2, 2))).astype(np.float32), dt=u.FP16)
stats = torch.zeros((148, 16), dtype=torch.int64, device='cuda')
lib = u._lib.lib()
lib.mmlf_debug_conv_stats.argtypes = [C.c_void_p]
REPS = int(os.environ.get('CONV_STATS_REPS', '3'))
for i in range(REPS):
    u.run_conv(x, cin_pad, cin_pad, w, n_pad, B, H, W, ctype, relu=True, ab=u.FP16, out_dt=u.FP16)
lib.mmlf_debug_conv_stats(C.c_void_p(stats.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
u.run_conv(x, cin_pad, cin_pad, w, n_pad, B, H, W, ctype, relu=True, ab=u.FP16, out_dt=u.FP16)
e1.record(); torch.cuda.synchronize()
lib.mmlf_debug_conv_stats(C.c_void_p(0))
s = stats.cpu().numpy().astype(np.float64)
lead, peer = s[0::2], s[1::2]
names = ['producer total', 'producer wait-empty', 'mma total', 'mma wait-full(TMA)', 'mma wait-tmem(epilogue)', 'epilogue total', 'epilogue wait-acc']
print(f'{B}x{H}x{W} {cin}->{cout} type {ctype}: {e0.elapsed_time(e1)*1e3:.1f} us; tiles/pair = {np.ceil(n_slots/256)/74:.2f}')
for i, n in enumerate(names):
    print(f'  {n:28s} leader {lead[:, i].mean():10.0f}  peer {peer[:, i].mean():10.0f} cycles')
print(f'  mma issue blocks {lead[:, 12].mean():10.0f}   commits {lead[:, 13].mean():10.0f} cycles')
print(f'  SM clock during the kernel: {lead[:, 0].mean() / lead[:, 7].mean() * 1e3:.0f} MHz (producer cycles / globaltimer ns)')
t0 = s[:, 8].min()
for nm, col in (('kernel entry', 8), ('main loop entry', 9), ('epilogue done', 10), ('kernel exit', 11)):
    print(f'  {nm:18s} first {(s[:, col].min() - t0) / 1e3:7.2f} us   last {(s[:, col].max() - t0) / 1e3:7.2f} us')
flops = 2.0 * n_slots * cout * 4 * cin
print(f'  {flops / (e0.elapsed_time(e1) * 1e-3) / 1e12:.1f} TFLOP/s algorithmic')
