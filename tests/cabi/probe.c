/* Plain-C consumer of the drop-in boundary (include/i2sdf_b200.h): what a cgo / JNI / N-API stub on the reference side would do.
 * Links nothing but libdl; prints one line per check for tests/test_host_cpu.py::test_c_abi_from_plain_c.
 *   argv[1] = path of libi2sdf_b200.so
 * No compute call is made: create() is expected to fail LOUDLY (negative status + message, no crash) on a host without an
 * sm_100 GPU and to succeed on a B200 (then the handle is queried and destroyed). */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "i2sdf_b200.h"

#define SYM(type, name) type name##_fn = (type)dlsym(lib, #name); if (!name##_fn) { printf("missing %s\n", #name); return 3; }

typedef int (*ver_t)(void);
typedef const char* (*err_t)(void);
typedef int (*create_t)(const i2sdf_desc*, int, i2sdf_handle**);
typedef int (*destroy_t)(i2sdf_handle*);
typedef int (*nlayers_t)(const i2sdf_handle*);
typedef size_t (*slot_t)(int64_t, int);
typedef int (*loss_t)(const i2sdf_loss_args*, void*);

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!lib) { printf("dlopen failed: %s\n", dlerror()); return 2; }
    SYM(ver_t, i2sdf_abi_version)
    SYM(err_t, i2sdf_last_error)
    SYM(create_t, i2sdf_create)
    SYM(destroy_t, i2sdf_destroy)
    SYM(nlayers_t, i2sdf_num_layers)
    SYM(slot_t, i2sdf_planes_slot_bytes)
    SYM(loss_t, i2sdf_loss_forward)
    printf("abi %d header %d\n", i2sdf_abi_version_fn(), I2SDF_ABI_VERSION);
    printf("slot256 %zu slot48 %zu\n", i2sdf_planes_slot_bytes_fn(1000, 256), i2sdf_planes_slot_bytes_fn(1000, 48));

    /* 1. a description the library must reject before touching the device */
    i2sdf_desc bad;
    memset(&bad, 0, sizeof bad);
    bad.abi_version = I2SDF_ABI_VERSION;
    bad.hidden = 128;
    i2sdf_handle* h = NULL;
    int rc = i2sdf_create_fn(&bad, 0, &h);
    printf("create_bad rc %d handle %s msg %s\n", rc, h ? "set" : "null", i2sdf_last_error_fn());

    /* 2. argument validation of a handle-less entry point (no launch happens: rgb is NULL) */
    i2sdf_loss_args la;
    memset(&la, 0, sizeof la);
    rc = i2sdf_loss_forward_fn(&la, NULL);
    printf("loss_null rc %d msg %s\n", rc, i2sdf_last_error_fn());

    /* 3. config/synthetic.yml: 9 SDF layers (skip at 4, PE 6), 5 radiance layers (PE 4), sampler 64 / 128 / 32, 10 / 5 iterations */
    static float u_up[128], u_final[64], t_init[128];
    static int32_t extra[5 * 32];
    for (int i = 0; i < 128; ++i) u_up[i] = t_init[i] = (float)i / 127.0f;
    for (int i = 0; i < 64; ++i) u_final[i] = (float)i / 63.0f;
    for (int k = 0; k < 5; ++k) for (int i = 0; i < 32; ++i) extra[k * 32 + i] = (int32_t)((128.0 * (k + 1) - 1.0) * i / 31.0);
    i2sdf_desc d;
    memset(&d, 0, sizeof d);
    d.abi_version = I2SDF_ABI_VERSION;
    d.hidden = 256; d.feature_size = 256;
    d.n_sdf_layers = 9; d.sdf_skip_layer = 4; d.multires_x = 6;
    d.n_color_layers = 5; d.multires_d = 4;
    d.n_light_layers = 0; d.light_hidden = 128;
    d.n_samples = 64; d.n_samples_eval = 128; d.n_samples_extra = 32; d.beta_iters = 10; d.max_total_iters = 5;
    d.near_ = 0.0f; d.far_ = 6.0f; d.eps = 0.1f; d.add_tiny = 1e-6f; d.beta_min = 1e-4f; d.lemma2_coeff = 2.6230f;
    d.u_up = u_up; d.u_final = u_final; d.t_init = t_init; d.extra_idx = extra;
    h = NULL;
    rc = i2sdf_create_fn(&d, 0, &h);
    if (rc == I2SDF_OK) {
        printf("create_ok rc 0 layers %d\n", i2sdf_num_layers_fn(h));
        printf("destroy rc %d\n", i2sdf_destroy_fn(h));
    } else {
        printf("create_nogpu rc %d handle %s msg %s\n", rc, h ? "set" : "null", i2sdf_last_error_fn());
    }
    dlclose(lib);
    return 0;
}
