/* A plain C99 client of the C ABI: proves include/qmps_b200.h is a valid C header and that the shared library can
 * be bound without C++ or Python (what a cgo / JNI / FFI stub would do).  Test infrastructure. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include "../../include/qmps_b200.h"

typedef const char* (*str_fn)(void);
typedef int (*fp_fn)(int, int, int64_t, const void*, int64_t, const void*, int, int, void*, void*, void*, void*, void*,
                     int32_t*, int, void*);

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 3; }
  str_fn version = (str_fn)dlsym(h, "qmps_version");
  str_fn last_error = (str_fn)dlsym(h, "qmps_last_error");
  fp_fn fixed_point = (fp_fn)dlsym(h, "qmps_fixed_point");
  if (!version || !last_error || !fixed_point) return 4;
  if (!strstr(version(), "sm_100a")) return 5;
  /* bad bond dimension: rejected before any CUDA call, message available */
  int rc = fixed_point(2, 17, 1, (const void*)1, 1, (const void*)1, 0, 0, 0, 0, 0, 0, 0, 0, QMPS_C128, 0);
  if (rc != QMPS_ERR_UNSUPPORTED || !strstr(last_error(), "D must be")) return 6;
  /* empty batch: success */
  if (fixed_point(2, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, QMPS_C128, 0) != 0) return 7;
  printf("c_client ok: %s\n", version());
  return 0;
}
