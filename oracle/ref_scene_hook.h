/*
 * ref_scene_hook.h -- force-included (nvcc -include) in front of the UNMODIFIED reference
 * src/Global{Float,Double}CUDAInOneWeekend/main.cu to capture the scene it uploads.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref/scene_dump_*).
 *
 * The reference builds its scene on the host and ships it with three cudaMemcpy calls
 * (GF main.cu:303-314: materials, spheres, world).  The macros below turn every CUDA runtime
 * call made before that point into a host stub, so the program runs on a machine without a GPU,
 * appends each host->device payload to $ORC_SCENE_DUMP as {uint64 size, bytes}, and exits after
 * the third copy -- before the first kernel launch.
 */
#ifndef REF_SCENE_HOOK_H
#define REF_SCENE_HOOK_H
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>

static inline cudaError_t orc_hook_malloc(void **p, size_t n) {
    *p = std::malloc(n ? n : 1);
    return cudaSuccess;
}
static inline cudaError_t orc_hook_memcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) {
    static int copies = 0;
    const char *path = std::getenv("ORC_SCENE_DUMP");
    if (path) {
        FILE *f = std::fopen(path, copies == 0 ? "wb" : "ab");
        if (!f) { std::perror("ORC_SCENE_DUMP"); std::exit(2); }
        uint64_t sz = n;
        std::fwrite(&sz, sizeof sz, 1, f);
        std::fwrite(src, 1, n, f);
        std::fclose(f);
    }
    std::memcpy(dst, src, n);
    if (++copies == 3) std::exit(0);
    return cudaSuccess;
}
#define cudaSetDevice(dev)            cudaSuccess
#define cudaEventCreate(ev)           cudaSuccess
#define cudaEventRecord(ev, stream)   cudaSuccess
#define cudaMallocManaged(p, n)       orc_hook_malloc((void **)(p), (n))
#define cudaMalloc(p, n)              orc_hook_malloc((void **)(p), (n))
#define cudaMemcpy(d, s, n, k)        orc_hook_memcpy((d), (s), (n), (k))
#endif
