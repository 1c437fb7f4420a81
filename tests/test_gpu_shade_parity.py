"""GPU parity for the SHADE stage, value by value (VERDICT r01 missing #7 / weak #1).

material::emitted, material::scatter, texture::value, perlin::noise and camera::get_ray are deterministic once the random
numbers they consume are given (reference: src/rendering/material.h:25-131, texture.h:16-154, perlin.h:21-96,
src/engine/camera.h:38-47). artb_shade_batch / artb_camera_rays run the device functions the render kernels run on
explicit inputs with EXPLICIT random numbers; the oracle (oracle/port: the reference's scatter / get_ray with the same xi
fed through the same closed-form samplers) must then agree:
  * emitted and attenuation (solid, checker, Perlin noise, image and barycentric-image texture values): bit-identical
    (image textures: a texel flip is allowed for u,v within an ulp of a texel border, < 0.2 % of hits)
  * scattered origin: bit-identical for surface hits; scattered direction and time: 1e-13 relative (sin/cos/cbrt of the
    CUDA and glibc math libraries differ in the last place)
  * scatter()'s boolean (metal: dot(scattered, normal) > 0; dielectric: reflect vs refract): identical
Rays: the unmodified reference's golden batches (all bounces; inside glass; inside media; unnormalised triangle normals
under metal for aliases 13/14)."""
import numpy as np
import pytest

from conftest import GOLDEN
import oracle_lib as O

pytestmark = pytest.mark.gpu

ALIASES = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14]
IMAGE_TEXTURED = {4, 8, 9}          # earth map / capsule map_Kd


@pytest.mark.parametrize("alias", ALIASES)
def test_emitted_and_scatter_match_oracle(artb, gpu_ctx, alias):
    g = np.load(GOLDEN / f"trace_{alias}.npz")
    rays = g["rays"]
    n = len(rays)
    mxi = g["xi"] if int(g["n_media"]) else None
    xi = np.random.default_rng(1000 + alias).uniform(0.0, 1.0, (n, 4))
    oracle = O.OracleScene(alias)
    ohit, osh = oracle.shade(rays, xi, medium_xi=mxi)
    scene = artb.Scene.build(alias)
    gpu_ctx.upload(scene)
    hit = gpu_ctx.trace_batch(rays, 0.001, float("inf"), mxi)
    h = ohit["hit"].astype(bool)
    assert np.array_equal(hit["hit"].astype(bool), h) and np.array_equal(hit["prim_id"][h], ohit["prim_id"][h])
    gsh = gpu_ctx.shade_batch(rays[h], hit["prim_id"][h], hit["t"][h], xi[h])
    o = osh[h]
    assert np.all(gsh["scattered"] >= 0)
    assert np.array_equal(gsh["material_type"], o["material_type"])
    assert np.array_equal(gsh["front_face"], o["front_face"])
    assert np.array_equal(gsh["scattered"], o["scattered"]), "scatter() must return the same boolean"
    assert np.array_equal(gsh["emitted"], o["emitted"]), "emitted radiance must be bit-identical"
    same_att = np.all(gsh["attenuation"] == o["attenuation"], axis=1)
    if alias in IMAGE_TEXTURED:
        assert np.mean(~same_att) < 2e-3, f"{(~same_att).sum()} attenuation values differ"
    else:
        assert same_att.all(), f"{(~same_att).sum()} of {len(o)} attenuation values differ from the oracle"
    surface = o["material_type"] != 4
    not_light = o["material_type"] != 3
    assert np.array_equal(gsh["origin"][surface & not_light], o["origin"][surface & not_light])
    if (~surface).any():
        assert np.max(np.abs(gsh["origin"][~surface] - o["origin"][~surface])) <= 1e-10
    assert np.array_equal(gsh["time"], o["time"])
    d_ref, d_gpu = o["direction"], gsh["direction"]
    scale = np.maximum(np.linalg.norm(d_ref, axis=1, keepdims=True), 1e-300)
    assert np.max(np.abs(d_gpu - d_ref) / scale, initial=0) <= 1e-13
    # the sample must have exercised what the scene has
    assert len(np.unique(o["material_type"])) >= 1 and h.sum() >= 50


def test_unknown_hit_ids_are_flagged(artb, gpu_ctx):
    gpu_ctx.upload(artb.Scene.build(6))
    rays = np.zeros((3, 7)); rays[:, 5] = 1
    out = gpu_ctx.shade_batch(rays, np.array([-1, 10 ** 6, 0]), np.ones(3), np.full((3, 4), 0.5))
    assert out["scattered"][0] == -1 and out["scattered"][1] == -1 and out["scattered"][2] >= 0


@pytest.mark.parametrize("alias", [1, 6, 8])
def test_camera_rays_match_oracle(artb, gpu_ctx, alias):
    """_stochastic_sample's (s, t) + camera::get_ray: alias 1 has a lens (aperture 0.1), all have the [0,1) shutter"""
    scene = artb.Scene.build(alias)
    w, h = 200, 120
    cam = scene.camera(w, h)
    rng = np.random.default_rng(alias)
    n = 4096
    st = rng.uniform(0, 1, (n, 2))
    xi = rng.uniform(0, 1, (n, 3))
    got = gpu_ctx.camera_rays(cam, st, xi)
    ref = O.camera_rays(cam.lookfrom, cam.lookat, cam.vup, cam.vfov, cam.aspect_ratio, cam.aperture, cam.focus_dist, cam.time0,
                        cam.time1, st, xi)
    assert np.array_equal(got[:, 6], ref[:, 6]), "ray time must be bit-identical"
    if cam.aperture == 0:
        assert np.array_equal(got, ref), "pinhole camera: origin and direction bit-identical"
    else:
        assert np.max(np.abs(got - ref)) <= 1e-13
