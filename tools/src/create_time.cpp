// Times dlopen(libb2pt.so) and b2pt_create separately (tools/time_phases.py sees 2-6 s in Context()).
#include <dlfcn.h>
#include <chrono>
#include <cstdio>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
    double t0 = now(), t;
    void *lib = dlopen(argv[1], RTLD_NOW);
    t = now(); printf("dlopen libb2pt.so              %8.1f ms (%s)\n", t - t0, lib ? "ok" : dlerror()); t0 = t;
    if (!lib) return 1;
    typedef int (*create_t)(void **, int);
    typedef void (*destroy_t)(void *);
    create_t create = (create_t)dlsym(lib, "b2pt_create");
    destroy_t destroy = (destroy_t)dlsym(lib, "b2pt_destroy");
    void *ctx = nullptr;
    int r = create(&ctx, 0);
    t = now(); printf("b2pt_create                    %8.1f ms (rc %d)\n", t - t0, r); t0 = t;
    void *ctx2 = nullptr;
    r = create(&ctx2, 0);
    t = now(); printf("second b2pt_create             %8.1f ms (rc %d)\n", t - t0, r); t0 = t;
    destroy(ctx2); destroy(ctx);
    t = now(); printf("destroy x2                     %8.1f ms\n", t - t0);
    return 0;
}
