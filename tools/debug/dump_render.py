import sys, numpy as np
sys.path.insert(0, '.')
import another_raytracer_b200 as A
alias, w, h, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ctx = A.Context(0, 1 << 20)
s = A.Scene.build(alias); ctx.upload(s)
rgb, st, acc, _ = ctx.render(s.camera(w, h), A.RenderOptions(w, h, n, seed=3), want_accum=True)
np.savez_compressed(f'gpurun_out/render_{alias}_{w}x{h}x{n}.npz', lin=(acc.astype(np.float64)/n).astype(np.float32), rgb=rgb)
print(st)
