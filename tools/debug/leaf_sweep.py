import sys, time
sys.path.insert(0, '.')
import another_raytracer_b200 as A
ctx = A.Context(0, 1 << 21)
for leaf in (1, 2, 3):
    for dev in (True, False):
        sc = A.Scene.build(12, mesh_as_soup=True, device_bvh=dev, max_leaf_prims=leaf)
        ctx.upload(sc)
        cam = sc.camera(1920, 1080)
        for k in range(2):
            rgb, st, _, _ = ctx.render(cam, A.RenderOptions(1920, 1080, 16))
        print(f"leaf {leaf} device_bvh {dev}: nodes {ctx.last_bvh_build['n_nodes'] if dev else sc.info['n_nodes']} {st['samples'] / st['elapsed_ms'] / 1e3:.1f} Msamples/s {st['rays'] / st['elapsed_ms'] / 1e3:.1f} Mrays/s", flush=True)
