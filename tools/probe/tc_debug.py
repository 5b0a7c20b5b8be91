"""GPU diagnostic: dump the raw conv1/conv2 accumulators of CTA (0,0) of the tcgen05 GuidanceNet kernel and compare
with a numpy emulation of the same linearised implicit GEMM."""
import ctypes as C, sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rt_octree_b200 import capi
g = np.load('tests/golden/guidance_net_ref.npz')
w = {k: g[k] for k in ('w1', 'b1', 'w2', 'b2')}
H, W = 24, 80
rs = np.random.default_rng(3); aux = rs.uniform(0, 1, (8, H, W)).astype(np.float32); aux[4:] = aux[:4] ** 2
L = capi.load()
net = capi.Denoiser(w); net.set_impl(0)
dbg = torch.full((1024 * 40 + 32768 + 4608,), -777.0, device='cuda')
L.rto_debug_tc_dump.argtypes = [C.c_void_p]
assert L.rto_debug_tc_dump(dbg.data_ptr()) == 0
a = torch.from_numpy(aux).cuda(); wm = torch.zeros((4, H, W), device='cuda'); gm = torch.zeros((4, H, W), device='cuda')
net.forward(a.data_ptr(), W, H, wm.data_ptr(), gm.data_ptr()); torch.cuda.synchronize()
d = dbg.cpu().numpy(); c1 = d[:1024 * 32].reshape(1024, 32); c2 = d[1024 * 32:1024 * 40].reshape(1024, 8)
mid_g = d[1024 * 40:1024 * 40 + 32768].reshape(4, 1024, 8); w2_g = d[1024 * 40 + 32768:].reshape(9, 4, 16, 8)
TW, TH, PW = 60, 12, 64; IN_PX = (TH + 4) * PW + 64; Q1 = PW + 1; Q2 = 2 * PW + 2
h = lambda v: v.astype(np.float16).astype(np.float32)
w1 = w['w1'].astype(np.float32); w2 = w['w2'].astype(np.float32); b1 = w['b1'].astype(np.float32)
inimg = np.zeros((IN_PX + 200, 8), np.float32)
for p in range(IN_PX):
    x, y = p & 63, p >> 6; gx, gy = x - 2, y - 2
    if y < TH + 4 and 0 <= gx < W and 0 <= gy < H: inimg[p] = h(aux[:, gy, gx])
e1 = np.zeros((1024, 32), np.float32); mid = np.zeros((1224, 32), np.float32)
for q in range(Q1, 961):
    acc = np.zeros(32, np.float32)
    for t in range(9): acc += w1[:, :, t // 3, t % 3] @ inimg[q + (t // 3 - 1) * PW + (t % 3 - 1)]
    e1[q] = acc; x, y = q & 63, q >> 6; gx, gy = x - 2, y - 2
    mid[q] = np.clip(h(h(acc) + b1), 0, 6) if (0 <= gx < W and 0 <= gy < H) else 0
e2 = np.zeros((1024, 8), np.float32)
for q in range(Q2, 898):
    acc = np.zeros(8, np.float32)
    for t in range(9): acc += w2[:, :, t // 3, t % 3] @ mid[q + (t // 3 - 1) * PW + (t % 3 - 1)]
    e2[q] = acc
d1 = np.abs(c1[Q1:961] - e1[Q1:961]); d2 = np.abs(c2[Q2:898] - e2[Q2:898])
print('conv1 raw acc: max err %.4g, rows wrong %d/%d ; untouched %d' % (d1.max(), (d1.max(1) > 1e-3).sum(), d1.shape[0], (c1[Q1:961] == -777).sum()))
print('conv2 raw acc: max err %.4g, rows wrong %d/%d ; untouched %d' % (d2.max(), (d2.max(1) > 2e-2).sum(), d2.shape[0], (c2[Q2:898] == -777).sum()))
bad1 = np.where(d1.max(1) > 1e-3)[0] + Q1; bad2 = np.where(d2.max(1) > 2e-2)[0] + Q2
print('conv1 bad q (first 40):', bad1[:40], ' bad channel histogram:', (d1 > 1e-3).sum(0))
print('conv2 bad q (first 40):', bad2[:40], ' bad channel histogram:', (d2 > 2e-2).sum(0))
if len(bad1): q = bad1[0]; print('q', q, 'gpu', c1[q, :8], 'ref', e1[q, :8])
if len(bad2): q = bad2[0]; print('q', q, 'gpu', c2[q], 'ref', e2[q])

mid_ref = mid[:1024].reshape(1024, 4, 8).transpose(1, 0, 2)
dm = np.abs(mid_g[:, Q1:961] - mid_ref[:, Q1:961])
print('mid smem vs ref: max', np.nanmax(dm), 'nan', np.isnan(mid_g[:, Q1:961]).sum(), 'wrong', (dm > 2e-3).sum(), '/', dm.size, 'untouched', (mid_g[:, Q1:961] == -777).sum())
w2_ref = np.zeros((9, 4, 16, 8), np.float32)
for t in range(9):
    for c in range(4):
        w2_ref[t, c, :8, :] = w2[:, c * 8:(c + 1) * 8, t // 3, t % 3]
print('w2 smem vs ref: max', np.abs(w2_g - w2_ref).max())
mg = np.zeros((1224, 32), np.float32); mg[:1024] = np.nan_to_num(mid_g.transpose(1, 0, 2).reshape(1024, 32), nan=1e9)
e2g = np.zeros((1024, 8), np.float32)
for q in range(Q2, 898):
    acc = np.zeros(8, np.float32)
    for t in range(9): acc += w2[:, :, t // 3, t % 3] @ mg[q + (t // 3 - 1) * PW + (t % 3 - 1)]
    e2g[q] = acc
print('conv2 acc vs numpy-on-GPU-mid: max', np.nanmax(np.abs(c2[Q2:898] - e2g[Q2:898])))
bw = np.argwhere(dm > 2e-3)
print('first wrong mid entries (plane, q-Q1, k):', bw[:12].tolist())
