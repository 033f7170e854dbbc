// Ordering of the compact hit list (cgbk_hit8) by (position, + strand first): one CUB radix sort
// over the upper 32 bits of the records (position and strand; the float32 score rides along in
// the lower half).  The hit list of a scan comes out in the order the warps found the sites;
// the host join with the gene regions (cgb/genome.py:419-445 restated as a search over sorted
// positions, cgb_b200/site_search.py) wants it sorted.  Library code on purpose: sorting is not
// the path's hot operation.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "cgbk_internal.cuh"

static int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

static int64_t cub_temp_bytes(int64_t n) {
    size_t bytes = 0;
    cub::DoubleBuffer<unsigned long long> db(nullptr, nullptr);
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, db, n, 32, 64, (cudaStream_t)0);
    return (int64_t)bytes;
}

extern "C" int64_t cgbk_sort_hits_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    return align256(8 * n) + align256(cub_temp_bytes(n)) + 256;
}

extern "C" int cgbk_sort_hits(uint64_t* d_hits8, int64_t n, void* d_tmp, int64_t tmp_bytes, void* stream) {
    if (n < 0) return cgbk_set_error(CGBK_ERR_INVALID, "negative count");
    if (n < 2) return CGBK_OK;
    if (!d_hits8 || !d_tmp || tmp_bytes < cgbk_sort_hits_workspace_bytes(n))
        return cgbk_set_error(CGBK_ERR_INVALID, "sort workspace of %lld bytes is smaller than the %lld required",
                              (long long)tmp_bytes, (long long)cgbk_sort_hits_workspace_bytes(n));
    ON_DEVICE(cgbk_device_of(d_hits8));
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* alt = (unsigned long long*)d_tmp;
    void* cub_tmp = (unsigned char*)d_tmp + align256(8 * n);
    size_t cub_bytes = (size_t)(tmp_bytes - align256(8 * n));
    cub::DoubleBuffer<unsigned long long> db((unsigned long long*)d_hits8, alt);
    cudaError_t e = cub::DeviceRadixSort::SortKeys(cub_tmp, cub_bytes, db, n, 32, 64, st);
    if (e != cudaSuccess) return cgbk_check_cuda(e, "radix sort of the hit list");
    if (db.Current() != (unsigned long long*)d_hits8) {
        e = cudaMemcpyAsync(d_hits8, db.Current(), 8 * (size_t)n, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cgbk_check_cuda(e, "copy of the sorted hit list");
    }
    return CGBK_OK;
}

// Exclusive prefix sum over the per-tile counts of the fused scan's site pass (n = tiles of one
// scan, at most a few hundred thousand).
long long cgbk_scan_tmp_bytes(long long n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int*)nullptr, (int*)nullptr, n, (cudaStream_t)0);
    return (long long)align256((int64_t)bytes) + 256;
}

cudaError_t cgbk_exclusive_sum(const int* in, int* out, long long n, void* tmp, long long tmp_bytes, cudaStream_t st) {
    size_t bytes = (size_t)tmp_bytes;
    return cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, st);
}
