"""Model behind the GEMM tile order (csrc/gemm.cu: tile_coords): average number of operand panels (A row-panels +
B row-panels, 256 rows x K each) that one wave of 74 CTA pairs keeps live in L2, for m-groups sweeping n (group > 0,
the first version) and n-groups sweeping m (group < 0), at the bench shapes. Fewer live panels = less DRAM traffic
(measured on the K = 28688 grad_input GEMM: 21.7 -> 18.3 panels in the model, 9.5 -> 8.5 GB and 2.76 -> 2.66 ms on a B200).
usage: python tools/tile_order_model.py"""


def coords(t, nm, nn, g):
    if g > 0:
        per = g * nn
        grp = t // per
        fm = grp * g
        gs = min(nm - fm, g)
        r = t - grp * per
        return fm + r % gs, r // gs
    g = -g
    per = g * nm
    grp = t // per
    fn = grp * g
    gs = min(nn - fn, g)
    r = t - grp * per
    return r // gs, fn + r % gs


def live_panels(nm, nn, g, clusters=74):
    tot = waves = 0
    for t0 in range(0, nm * nn, clusters):
        ms, ns = set(), set()
        for t in range(t0, min(t0 + clusters, nm * nn)):
            a, b = coords(t, nm, nn, g)
            ms.add(a)
            ns.add(b)
        tot += len(ms) + len(ns)
        waves += 1
    return tot / waves


if __name__ == "__main__":
    shapes = {"grad_input N=4096 (wqkv, wo, w1|w3)": (16384, 4096), "w1 / w3 forward, w2 grad_input N=14336": (16384, 14336),
              "wk / wv forward N=1024": (16384, 1024), "LM head dx chunk 8192 x 4096": (8192, 4096),
              "LM head logits chunk 8192 x 128256": (8192, 128256), "speech batch M=14048, N=4096": (14048, 4096)}
    print(f"{'shape':44s} {'tiles':>9s}  m-groups(8)  n-groups(8)  launcher picks   (2*sqrt(74) = 17.2 is the floor)")
    for name, (M, N) in shapes.items():
        nm, nn = (M + 255) // 256, (N + 255) // 256
        pick = -8 if nm >= nn else 8
        print(f"{name:44s} {nm:4d}x{nn:<4d}  {live_panels(nm, nn, 8):11.2f}  {live_panels(nm, nn, -8):11.2f}  {pick:+d}")
