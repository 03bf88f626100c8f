"""CPU model (numpy) of what the tile bitmaps can skip on a d = 64 Gaussian mixture under the two-means partition of
gpu_build.cu: tile purity, seeds from the home bucket under both routings (nearer centroid / split planes), and the share of
tiles a single query, a warp of 32, a subtile of 128 and a CTA group of 512 sorted queries still need.  It guided the design
(DESIGN.md 4.5): run as  python scripts/partition_sim.py 262144 256  (points, clusters)."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from petal_neighbors_b200 import synth
n, d, nc = int(sys.argv[1]), 64, int(sys.argv[2]); sigma = 0.05
pts = synth.gaussian_mixture(n, d, 1, n_centers=nc, sigma=sigma).astype(np.float32)
Q = synth.gaussian_mixture(n, d, 2, n_centers=nc, sigma=sigma).astype(np.float32)
def amid(s, e):
    T = (e - s + 127)//128
    return s + 128*(T//2) if T >= 2 else (s+e)//2
def maxlen(L):
    sg=[(0,n)]
    for l in range(L): sg=[x for (s,e) in sg for x in ((s,amid(s,e)),(amid(s,e),e))]
    return max(e-s for s,e in sg)
L = 0
while maxlen(L) > 256: L += 1
idx = np.arange(n)
def direction(X):
    m = 256; L_ = len(X)
    S = X if L_ <= m else X[(np.arange(m) * L_) // m]
    a = S[0]; b = S[((S-a)**2).sum(1).argmax()]; c = S[((S-b)**2).sum(1).argmax()]
    for it in range(4):
        mk = ((S-b)**2).sum(1) <= ((S-c)**2).sum(1)
        if mk.all() or (~mk).all(): break
        b = S[mk].mean(0); c = S[~mk].mean(0)
    return c-b
segs_at = [[(0, n)]]; W = []; TH = []
for l in range(L + 1):
    new = []; Wl = []; Tl = []
    for (s, e) in segs_at[-1]:
        mid = amid(s, e)
        if e - s >= 2:
            X = pts[idx[s:e]]; w = direction(X); key = X @ w
            o = np.argpartition(key, mid - s); idx[s:e] = idx[s:e][o]
            Wl.append(w); Tl.append(key[o[mid - s]])
        else:
            Wl.append(np.zeros(d, np.float32)); Tl.append(0.0)
        new += [(s, mid), (mid, e)]
    segs_at.append(new); W.append(np.array(Wl)); TH.append(np.array(Tl))
P = pts[idx]
bk = segs_at[L]   # buckets
cen_at = [np.array([P[s:e].mean(0) if e > s else np.zeros(d) for (s, e) in lv]) for lv in segs_at[:L+1]]
# routing
def route_centroid(Q):
    node = np.zeros(len(Q), np.int64)
    for l in range(1, L+1):
        c1 = cen_at[l][2*node]; c2 = cen_at[l][2*node+1]
        right = ((Q-c1)**2).sum(1) > ((Q-c2)**2).sum(1)
        node = 2*node + right
    return node
def route_proj(Q):
    node = np.zeros(len(Q), np.int64)
    for l in range(L):
        key = (Q * W[l][node]).sum(1)
        right = key >= TH[l][node]
        node = 2*node + right
    return node
nt = (n+127)//128
tc = np.array([P[t*128:(t+1)*128].mean(0) for t in range(nt)]); tr = np.array([np.sqrt(((P[t*128:(t+1)*128]-tc[t])**2).sum(1)).max() for t in range(nt)])
print('L', L, 'tiles', nt, 'tile radius median %.3f p75 %.3f p90 %.3f' % (np.median(tr), np.percentile(tr,75), np.percentile(tr,90)))
def seeds(home):
    sd = np.empty(len(Q), np.float32)
    order = np.argsort(home, kind='stable')
    hs = home[order]; bounds = np.searchsorted(hs, np.arange(len(bk)+1))
    for b in range(len(bk)):
        qi = order[bounds[b]:bounds[b+1]]
        if len(qi) == 0: continue
        s, e = bk[b]; X = P[s:e]
        d2 = (Q[qi]**2).sum(1)[:,None] + (X**2).sum(1)[None,:] - 2*Q[qi]@X.T
        sd[qi] = np.sqrt(np.maximum(d2.min(1), 0))
    return sd, order
tcn = (tc**2).sum(1)
bcen = np.array([P[s:e].mean(0) for (s, e) in bk]); brad = np.array([np.sqrt(((P[s:e]-bcen[i])**2).sum(1)).max() for i, (s, e) in enumerate(bk)])
def analyse(name, home, extra_seed=None, stray_factor=None):
    sd, order = seeds(home)
    if extra_seed is not None: sd = np.minimum(sd, extra_seed)
    if stray_factor is not None:
        dq = np.sqrt(((Q - bcen[home])**2).sum(1))
        stray = dq > stray_factor * brad[home]
        print('   strays: %.4f of the queries' % stray.mean())
        key = np.where(stray, len(bk), home)
        order = np.argsort(key, kind='stable')
    single = []; uni512 = []; uni128 = []; uni32=[]
    for g0 in range(0, len(Q), 512):
        qi = order[g0:g0+512]
        d2 = (Q[qi]**2).sum(1)[:,None] + tcn[None,:] - 2*Q[qi]@tc.T
        lb = np.sqrt(np.maximum(d2,0)) - tr[None,:]
        need = lb <= sd[qi][:,None]
        single.append(need.mean()); uni512.append(need.any(0).mean())
        uni128.append(np.mean([need[i:i+128].any(0).mean() for i in range(0, len(qi), 128)]))
        uni32.append(np.mean([need[i:i+32].any(0).mean() for i in range(0, len(qi), 32)]))
    print('%s: single %.4f  warp32 %.4f  sub128 %.4f  cta512 %.4f   seed p50 %.3f p99 %.3f max %.3f' % (name, np.mean(single), np.mean(uni32), np.mean(uni128), np.mean(uni512), np.median(sd), np.percentile(sd,99), sd.max()))
    return sd
hc = route_centroid(Q); hp = route_proj(Q)
print('routes agree %.4f' % (hc == hp).mean())
s1 = analyse('centroid routing', hc)
s2 = analyse('projection routing', hp)
analyse('projection routing + min with centroid-route seed', hp, s1)

for f in (1.5, 2.0):
    analyse('projection routing, strays (factor %.1f) grouped at the end' % f, hp, None, f)
