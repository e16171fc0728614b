// Cluster TRSV kernel (trsv_cluster.cuh), double arithmetic: all storage types,
// both triangles, both diagonal kinds, three row alignments.
#include "trsv_cluster.cuh"

namespace accblas {

int trsv_cluster_f64(Handle* h, int st, bool upper, bool unit, int vw,
                       std::int64_t n, const void* A, std::int64_t lda, void* x,
                       std::int64_t incx, void* xs, unsigned* ticket,
                       long long* trace, cudaStream_t stream)
{
    return trsv_cluster_ar<double>(h, st, upper, unit, vw, n, A, lda, x, incx, xs,
                                 ticket, trace, stream);
}

}  // namespace accblas
