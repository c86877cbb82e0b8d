"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libref_oracle.so = the reference's unmodified
sources compiled against oracle/eigen_shim).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

Inputs are seeded; outputs are what the reference's own code returns for them (Triangle::getIntersection,
Scene::intersect, Material::eval/pdf/fresnel/refract/sample, Scene::sampleLight/sampleEnv, the camera rays of
Renderer.cpp:44-76 and Scene::castRay per sample on the Philox sample streams).  The scenes are assembled from
assets/models/*.b2m, so the vectors can be replayed on machines without the reference (the GPU box)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import scenes  # noqa: E402
import support as S  # noqa: E402
from gen import adversarial_triangle_cases, bsdf_inputs, box_cases, sphere_cases, uniforms  # noqa: E402

b2pt = S.b2pt


def primitives():
    v, o, d = adversarial_triangle_cases(np.random.RandomState(101), 6000)
    hit, t = S.ref_tri(v, o, d)
    b6, bo, bd = box_cases(np.random.RandomState(102), 6000)
    c4, so, sd = sphere_cases(np.random.RandomState(103), 4000)
    shit, st, sco, snn = S.ref_sphere(c4, so, sd)
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), tri_v=v, tri_o=o, tri_d=d, tri_hit=hit, tri_t=t, box_b=b6, box_o=bo, box_d=bd,
                        box_hit=S.ref_box(b6, bo, bd), sph_c=c4, sph_o=so, sph_d=sd, sph_hit=shit, sph_t=st, sph_p=sco, sph_n=snn)


def bsdf():
    sc, _ = scenes.two_triangle_scene()
    ref = S.Ref(sc)
    rng = np.random.RandomState(104)
    n = 1500
    wi, wo, nrm, wl, uv, rf = bsdf_inputs(rng, n)
    u2 = uniforms(rng, n, 2)
    out = dict(wi=wi, wo=wo, n=nrm, wl=wl, uv=uv, rf=rf, u2=u2, reflect=ref.reflect(0, wi, nrm))
    for mat, name in enumerate(b2pt.NAMED_MATERIALS):
        out[f"eval_{mat}"] = ref.bsdf_eval(mat, wi, wo, nrm, wl, uv, rf)
        out[f"pdf_{mat}"] = ref.bsdf_pdf(mat, wi, wo, nrm, wl, rf)
        out[f"fresnel_{mat}"] = ref.fresnel(mat, wi, nrm, wl)
        out[f"refract_{mat}"] = ref.refract(mat, wi, nrm, wl)
        out[f"sample_{mat}"] = ref.material_sample(mat, wo, nrm, u2)
    np.savez_compressed(os.path.join(HERE, "bsdf.npz"), **out)
    ref.close(); sc.close()


def scene_vectors(name, sc, env):
    ref = S.Ref(sc, env)
    cam = sc.camera
    o, d, (p, ws, dist, u4) = scenes.ray_batch(ref, sc, n_pixels=900, samples=1, seed=105)
    prim, t, co, nn, uv = ref.intersect(o, d)
    sprim, st, *_ = ref.intersect(p, ws)
    visible = ((sprim >= 0) & (np.abs(st - dist.astype(np.float64)) < np.float64(np.float32(1e-4)))).astype(np.int32)
    u4 = uniforms(np.random.RandomState(106), 2000, 4)
    lp, ln, le, lpdf = ref.sample_light(u4)
    ed = np.random.RandomState(107).normal(size=(2000, 3)).astype(np.float32)
    px = np.random.RandomState(108).choice(cam.width * cam.height, 250, replace=False).astype(np.int32)
    co_, cd_ = ref.camera_rays(px, 2, 3)
    rad = ref.render_samples(px, 0, 4)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), ray_o=o, ray_d=d, prim=prim, t=t, coords=co, normal=nn, sh_o=p, sh_d=ws, sh_dist=dist,
                        sh_visible=visible, u4=u4, light_p=lp, light_n=ln, light_e=le, light_pdf=lpdf, env_d=ed, env_rgb=ref.sample_env(ed),
                        pixels=px, cam_o=co_, cam_d=cd_, radiance=rad, seed=np.uint64(S.SEED), width=cam.width, height=cam.height)
    ref.close()


def golden_scenes():
    return {"cornell": lambda: scenes.cornell(64, 64), "chess_sky_dof": lambda: scenes.chess(96, 54, dof=True, sky=True),
            "chess_dark": lambda: scenes.chess(96, 54, dof=False, sky=False)}


if __name__ == "__main__":
    S.ensure_built()
    primitives()
    bsdf()
    for name, make in golden_scenes().items():
        sc, env = make()
        scene_vectors(name, sc, env)
        sc.close()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
