/* The C-ABI boundary used from plain C: include/mxdet.h must compile as C99, the library must load without
 * Python / torch, and host-only entry points (version, struct size, workspace queries, argument validation) must
 * answer.  No kernel is launched: this runs on the CPU box.
 *   gcc -std=c99 -Wall -Wextra -pedantic -I include tests/c/abi_smoke.c -ldl -o abi_smoke && ./abi_smoke <path to .so> */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include "mxdet.h"

typedef int (*version_fn)(void);
typedef int (*sizeof_fn)(void);
typedef const char* (*err_fn)(void);
typedef size_t (*nms_ws_fn)(int, int);
typedef size_t (*roi_ws_fn)(int, int, int, int, const int*, const int*, int, int, int);
typedef size_t (*rpn_ws_fn)(const mxd_rpn_config*, int);
typedef int (*nms_fn)(const DLTensor*, const DLTensor*, const DLTensor*, DLTensor*, DLTensor*, float, float, int, float,
                      int, int, void*, size_t, void*);

#define LOAD(T, name) T name##_p; *(void**)(&name##_p) = dlsym(h, #name); if (!name##_p) { fprintf(stderr, "missing %s\n", #name); return 2; }

int main(int argc, char** argv) {
  void* h;
  int64_t shape[2] = {4, 4};
  DLTensor cpu;
  mxd_rpn_config cfg;
  int fh[4] = {200, 100, 50, 25}, fw[4] = {336, 168, 84, 42};
  size_t a, b, c;
  int rc;
  if (argc < 2) { fprintf(stderr, "usage: abi_smoke libmxdet_sm100.so\n"); return 2; }
  h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  {
    LOAD(version_fn, mxd_version)
    LOAD(sizeof_fn, mxd_sizeof_rpn_config)
    LOAD(err_fn, mxd_last_error)
    LOAD(nms_ws_fn, mxd_nms_workspace_bytes)
    LOAD(roi_ws_fn, mxd_roi_align_workspace_bytes)
    LOAD(rpn_ws_fn, mxd_rpn_proposals_workspace_bytes)
    LOAD(nms_fn, mxd_nms)
    if (mxd_version_p() < 100) { fprintf(stderr, "bad version\n"); return 1; }
    if (mxd_sizeof_rpn_config_p() != (int)sizeof(mxd_rpn_config)) { fprintf(stderr, "mxd_rpn_config layout differs\n"); return 1; }
    a = mxd_nms_workspace_bytes_p(2000, -1);
    b = mxd_roi_align_workspace_bytes_p(4096, 8, 256, 4, fh, fw, 7, 7, 2);
    memset(&cfg, 0, sizeof(cfg));
    cfg.num_levels = 1; cfg.feat_h[0] = 50; cfg.feat_w[0] = 68; cfg.stride[0] = 16.0f; cfg.num_base = 3;
    cfg.nms_pre = 2000; cfg.nms_post = 1000; cfg.max_num = 1000; cfg.nms_thr = 0.7f; cfg.delta = 1.0f;
    cfg.stds[0] = cfg.stds[1] = cfg.stds[2] = cfg.stds[3] = 1.0f; cfg.wh_ratio_clip = 16.0 / 1000.0;
    c = mxd_rpn_proposals_workspace_bytes_p(&cfg, 2);
    if (a == 0 || b == 0 || c == 0) { fprintf(stderr, "workspace queries: %lu %lu %lu\n", (unsigned long)a, (unsigned long)b, (unsigned long)c); return 1; }
    /* a CPU tensor is refused with MXD_ENOTSUP and a message - there is no CPU fallback */
    memset(&cpu, 0, sizeof(cpu));
    cpu.data = shape; cpu.device.device_type = kDLCPU; cpu.ndim = 2; cpu.dtype.code = kDLFloat; cpu.dtype.bits = 32;
    cpu.dtype.lanes = 1; cpu.shape = shape;
    rc = mxd_nms_p(&cpu, &cpu, NULL, &cpu, &cpu, 0.5f, 0.0f, -1, 0.0f, 1, -1, NULL, 0, NULL);
    if (rc != MXD_ENOTSUP || strlen(mxd_last_error_p()) == 0) { fprintf(stderr, "CPU tensor: rc %d\n", rc); return 1; }
    printf("abi ok: version %d, sizeof(mxd_rpn_config) %d, workspaces %lu / %lu / %lu bytes, CPU tensor refused: %s\n",
           mxd_version_p(), mxd_sizeof_rpn_config_p(), (unsigned long)a, (unsigned long)b, (unsigned long)c, mxd_last_error_p());
  }
  dlclose(h);
  return 0;
}
