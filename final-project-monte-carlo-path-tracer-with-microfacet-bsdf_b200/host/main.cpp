// host/main.cpp — ./RayTracing: the reference's program (src/main.cpp) with its hot path on the GPU.
//
// Same user interface as the reference: run from a directory that holds conf.json and sits beside
// ../models (main.cpp:25-26,147), no required arguments, the same stdout lines (" - Generating BVH...",
// "SPP: n", the progress bar, "Writing image to ...", "Rendering finished in H:M:S.ms") and the same
// RGBA8 PNG (gamma 0.45, clamp, truncation; Renderer.cpp:93-105).  Compile with -DDEMO (make DEMO=1) for
// the Cornell box of main.cpp:99-129, exactly like the reference's compile-time switch.
//
// Where the reference calls scene.buildBVH() and r.Render(scene) (main.cpp:330-333) this program calls
// b2pt_host_scene_build -> b2pt_upload_scene -> b2pt_render: the C ABI of include/b2pt.h.
//
// Optional arguments (the reference ignores argv; none is needed):
//   --demo                 DEMO scene without recompiling      --spp N, --width W, --height H   overrides
//   --conf PATH            another conf.json                   --device D                        CUDA device
//   --fix-ndir --fix-quality --fix-diamond --fix-output        opt-in corrections (include/b2pt_host.h)
//   --ndir N               next-event samples per vertex       --seed S                          sample-stream key
//   --chunk N              samples per pixel per b2pt_render call (progress granularity; default 256: the thin tail of deep paths at the end of a call costs ~2 ms)
//   --gpus N               split the samples over N GPUs of this box (one NCCL reduce of the frame per call)
//   --host-tonemap         tone map on the host like the reference (default: on the device, byte-identical)
//   --checkpoint FILE      after every chunk, save the fp32 accumulation + the samples done (the reference has no checkpoints:
//                          a 2048-spp run that dies after two hours starts over); --resume FILE continues such a run with the
//                          same sample streams, --stop-after N ends after N samples per pixel (still writes the PNG)
//   --preview-every N      also rewrite the PNG from the partial frame every N chunks (scaled by spp / samples done)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "b2pt.h"
#include "b2pt_host.h"

static void update_progress(float progress) {  // UpdateProgress, src/global.hpp:55-70
    int barWidth = 70;
    std::cout << "[";
    int pos = (int)(barWidth * progress);
    for (int i = 0; i < barWidth; ++i) {
        if (i < pos) std::cout << "=";
        else if (i == pos) std::cout << ">";
        else std::cout << " ";
    }
    std::cout << "] " << int(progress * 100.0) << " %\r";
    std::cout.flush();
}

int main(int argc, char **argv) {
#ifdef DEMO
    bool demo = true;
#else
    bool demo = false;
#endif
    std::string conf = "conf.json", run_dir = ".";
    int spp = 0, width = 0, height = 0, device = 0, fix = 0, ndir = 0, chunk = 256, gpus = 1;
    bool host_tonemap = false;
    std::string checkpoint, resume;
    int stop_after = 0, preview_every = 0;
    unsigned long long seed = 0x5EED0001ull;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : "0"; };
        if (a == "--demo") demo = true;
        else if (a == "--conf") conf = next();
        else if (a == "--spp") spp = std::atoi(next());
        else if (a == "--width") width = std::atoi(next());
        else if (a == "--height") height = std::atoi(next());
        else if (a == "--device") device = std::atoi(next());
        else if (a == "--ndir") ndir = std::atoi(next());
        else if (a == "--seed") seed = std::strtoull(next(), nullptr, 0);
        else if (a == "--chunk") chunk = std::max(1, std::atoi(next()));
        else if (a == "--host-tonemap") host_tonemap = true;
        else if (a == "--checkpoint") checkpoint = next();
        else if (a == "--resume") resume = next();
        else if (a == "--stop-after") stop_after = std::max(0, std::atoi(next()));
        else if (a == "--preview-every") preview_every = std::max(0, std::atoi(next()));
        else if (a == "--gpus") gpus = std::max(1, std::atoi(next()));
        else if (a == "--fix-ndir") fix |= B2PT_HOST_FIX_DIRECT_LIGHT_SAMPLE;
        else if (a == "--fix-quality") fix |= B2PT_HOST_FIX_MODEL_QUALITY;
        else if (a == "--fix-diamond") fix |= B2PT_HOST_FIX_ADD_DIAMOND;
        else if (a == "--fix-output") fix |= B2PT_HOST_FIX_OUTPUT_PATH;
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    {
        size_t s = conf.find_last_of('/');
        if (s != std::string::npos) run_dir = conf.substr(0, s);
    }
    if (const char *ad = std::getenv("B2PT_ASSET_DIR")) b2pt_host_set_asset_dir(ad);

    // Creating the CUDA contexts takes 1-2 s (driver initialisation + context: profiles/r02g_ctx_time.txt), more than everything the
    // host does before it needs them (scene assembly 0.1-0.4 s, trees 0.2 s): it runs on its own thread meanwhile.
    std::vector<b2pt_ctx *> ctxs(gpus, nullptr);
    std::vector<int> create_rc(gpus, B2PT_OK);
    std::string create_err;
    std::thread creator([&]() {
        for (int g = 0; g < gpus; ++g) {
            create_rc[g] = b2pt_create(&ctxs[g], device + g);
            if (create_rc[g] != B2PT_OK) { create_err = b2pt_last_error(nullptr); return; }
        }
    });
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{creator};

    b2pt_host_scene *scene = demo ? b2pt_host_scene_demo((run_dir + "/../models").c_str(), width, height)
                                  : b2pt_host_scene_from_conf(conf.c_str(), run_dir.c_str(), fix);
    if (!scene) { std::fprintf(stderr, "scene assembly failed: %s\n", b2pt_host_last_error()); return 1; }
    if (!demo && (width > 0 || height > 0)) {
        const b2pt_camera *c = b2pt_host_scene_camera(scene);
        b2pt_host_set_resolution(scene, width > 0 ? width : c->width, height > 0 ? height : c->height);
    }
    b2pt_host_set_render(scene, spp, -1.f, -1, ndir);

    std::printf(" - Generating BVH...\n\n");  // Scene::buildBVH, src/Scene.cpp:15
    if (b2pt_host_scene_build(scene) != 0) { std::fprintf(stderr, "BVH build failed: %s\n", b2pt_host_last_error()); return 1; }

    creator.join();
    for (int g = 0; g < gpus; ++g) {
        if (create_rc[g] != B2PT_OK || !ctxs[g]) { std::fprintf(stderr, "b2pt_create: %s\n", create_err.c_str()); return 1; }
        if (b2pt_upload_scene(ctxs[g], b2pt_host_scene_desc(scene)) != B2PT_OK) { std::fprintf(stderr, "b2pt_upload_scene: %s\n", b2pt_last_error(ctxs[g])); return 1; }
    }
    b2pt_ctx *ctx = ctxs[0];

    const b2pt_camera *cam = b2pt_host_scene_camera(scene);
    const int total_spp = b2pt_host_scene_spp(scene);
    const std::string path = b2pt_host_scene_output_path(scene);
    std::vector<float> framebuffer((size_t)cam->width * cam->height * 3, 0.f);

    auto start = std::chrono::system_clock::now();
    std::cout << "SPP: " << total_spp << "\n";
    unsigned long long rays = 0;
    double gpu_ms = 0;
    // checkpoint file: magic, width, height, spp_total, samples done, seed, a hash of everything else the samples depend on
    // (geometry, materials, lights, environment, camera, rr / shadow / light-sample settings), then the fp32 accumulation
    // (sum_k rgb_k / spp_total).  A resume with another conf.json, --ndir, --fix-* or --demo is refused instead of mixing
    // samples of two scenes into one frame.
    struct CkptHeader { char magic[8]; int32_t width, height, spp_total, done; uint64_t seed, scene_hash; };
    const uint64_t scene_hash = [&]() {
        uint64_t hsh = 1469598103934665603ull;  // FNV-1a
        auto mix = [&](const void *ptr, size_t bytes) {
            const unsigned char *b = (const unsigned char *)ptr;
            for (size_t i = 0; i < bytes; ++i) { hsh ^= b[i]; hsh *= 1099511628211ull; }
        };
        const b2pt_scene_desc *d = b2pt_host_scene_desc(scene);
        mix(&d->n_nodes, 4); mix(d->nodes, sizeof(b2pt_node) * (size_t)d->n_nodes);
        mix(&d->n_prims, 4);
        mix(d->prim_v0, 16 * (size_t)d->n_prims); mix(d->prim_e1, 16 * (size_t)d->n_prims); mix(d->prim_e2, 16 * (size_t)d->n_prims);
        mix(d->prim_uv, 24 * (size_t)d->n_prims); mix(d->prim_material, 4 * (size_t)d->n_prims); mix(d->prim_kind, 4 * (size_t)d->n_prims);
        for (uint32_t i = 0; i < d->n_materials; ++i) {  // field by field: the struct has a padding word
            const b2pt_material &m = d->materials[i];
            mix(&m.type, 4); mix(m.emission, 12); mix(&m.ior_a, 4); mix(&m.ior_b, 4); mix(&m.roughness, 4); mix(m.base_reflectance, 12); mix(&m.textured, 4);
        }
        mix(&d->n_lights, 4); mix(d->light_root, 4 * (size_t)d->n_lights); mix(d->light_material, 4 * (size_t)d->n_lights);
        mix(&d->use_env_map, 4); mix(&d->env_width, 4); mix(&d->env_height, 4);
        if (d->use_env_map) mix(d->env_rgb, 12 * (size_t)d->env_width * d->env_height);
        mix(d->background, 12); mix(&d->rr_rate, 4); mix(&d->enable_shadow, 4); mix(&d->n_dir_sample, 4);
        mix(&cam->width, 4); mix(&cam->height, 4); mix(cam->position, 12); mix(cam->orientation, 36); mix(&cam->scale, 4); mix(&cam->aspect, 4);
        mix(&cam->use_dof, 4); mix(&cam->focal_distance, 4); mix(&cam->aperture_radius, 4);
        return hsh;
    }();
    int first_sample = 0;
    if (!resume.empty()) {
        FILE *f = std::fopen(resume.c_str(), "rb");
        CkptHeader h{};
        if (!f || std::fread(&h, sizeof h, 1, f) != 1 || std::memcmp(h.magic, "B2PTCKP2", 8) != 0 || h.scene_hash != scene_hash || h.width != cam->width || h.height != cam->height ||
            h.spp_total != total_spp || h.seed != seed || h.done < 0 || h.done > total_spp ||
            std::fread(framebuffer.data(), sizeof(float), framebuffer.size(), f) != framebuffer.size()) {
            std::fprintf(stderr, "cannot resume from %s: missing, truncated, or written for another scene / configuration / frame size / spp / seed\n", resume.c_str());
            return 1;
        }
        std::fclose(f);
        first_sample = h.done;
    }
    auto save_checkpoint = [&](int done) {
        if (checkpoint.empty()) return true;
        const std::string tmp = checkpoint + ".tmp";
        FILE *f = std::fopen(tmp.c_str(), "wb");
        CkptHeader h{{'B', '2', 'P', 'T', 'C', 'K', 'P', '2'}, cam->width, cam->height, total_spp, done, seed, scene_hash};
        bool ok = f && std::fwrite(&h, sizeof h, 1, f) == 1 && std::fwrite(framebuffer.data(), sizeof(float), framebuffer.size(), f) == framebuffer.size();
        if (f) ok = (std::fclose(f) == 0) && ok;
        ok = ok && std::rename(tmp.c_str(), checkpoint.c_str()) == 0;  // a crash never leaves a half-written checkpoint behind
        if (!ok) std::fprintf(stderr, "\ncannot write checkpoint %s\n", checkpoint.c_str());
        return ok;
    };
    std::vector<unsigned char> raw((size_t)4 * cam->width * cam->height);
    auto write_image = [&](int done) {  // Renderer.cpp:93-109 on the frame accumulated so far
        if (done == total_spp && !host_tonemap && first_sample == 0 && b2pt_tonemap_rgba8(ctx, nullptr, cam->width * cam->height, raw.data()) == B2PT_OK) {
            // the whole frame is still resident on the device: 4 instead of 12 bytes per pixel come back, and the host is
            // spared 6 M double pow calls; byte-identical to the host loop
        } else if (done == total_spp) {
            b2pt_host_tonemap_rgba8(framebuffer.data(), cam->width * cam->height, raw.data());
        } else {  // partial frame: the accumulation is divided by spp_total, show it divided by the samples done
            std::vector<float> scaled(framebuffer.size());
            const float k = (float)total_spp / (float)std::max(done, 1);
            for (size_t i = 0; i < scaled.size(); ++i) scaled[i] = framebuffer[i] * k;
            b2pt_host_tonemap_rgba8(scaled.data(), cam->width * cam->height, raw.data());
        }
        return b2pt_host_write_png_rgba8(path.c_str(), raw.data(), cam->width, cam->height) == 0;
    };
    const int last_sample = stop_after > 0 ? std::min(total_spp, stop_after) : total_spp;
    int done = first_sample, chunks = 0;
    for (int s0 = first_sample; s0 < last_sample; s0 += chunk) {
        b2pt_render_params p{};
        p.spp_total = total_spp; p.sample_begin = s0; p.sample_count = std::min(chunk, last_sample - s0);
        p.seed = seed;
        // short jobs: a 16 Mi ray queue (27 GB) instead of the library's 48 Mi (82 GB) — allocating and releasing the larger one
        // costs more wall time than its 4 % buys unless the render runs for several seconds
        if ((double)cam->width * cam->height * total_spp < (double)(1ull << 31)) p.max_wave_bundles = 16 << 20;
        if (s0 == 0) p.flags |= B2PT_FLAG_FRESH_FRAME;  // first chunk starts the frame, later chunks accumulate
        b2pt_stats st{};
        if (b2pt_group_render(ctxs.data(), gpus, cam, &p, framebuffer.data(), &st) != B2PT_OK) { std::fprintf(stderr, "\nb2pt_render: %s\n", b2pt_last_error(ctx)); return 1; }
        rays += st.rays_reference;
        gpu_ms += st.gpu_ms;
        done = s0 + p.sample_count;
        ++chunks;
        if (!save_checkpoint(done)) return 1;
        if (preview_every > 0 && chunks % preview_every == 0 && done < last_sample) write_image(done);
        update_progress((float)done / (float)total_spp);
    }
    update_progress(1.f);
    std::cout << std::endl;
    std::cout << "Writing image to " << path << std::endl;
    if (!write_image(done)) std::cerr << "Error when writing image : " << b2pt_host_last_error() << std::endl;
    auto stop = std::chrono::system_clock::now();

    using Milli = std::chrono::milliseconds;
    long long ms = std::chrono::duration_cast<Milli>(stop - start).count();
    std::cout << "Rendering finished in " << ms / 3600000 << ":" << (ms / 60000) % 60 << ":" << (ms / 1000) % 60 << "." << ms % 1000 << std::endl;
    std::fprintf(stderr, "[b2pt] %.1f Mrays/s on the GPU (%llu rays, %.1f ms of device time)\n", gpu_ms > 0 ? rays / gpu_ms / 1e3 : 0.0, rays, gpu_ms);

    for (b2pt_ctx *c : ctxs) b2pt_destroy(c);
    b2pt_host_scene_free(scene);
    return 0;
}
