// Explicit instantiations of nuts_kernel for resident-operand layout TOEP = 2 (see nuts_kernel.cuh).
#include "nuts_kernel.cuh"

template __global__ void nuts_kernel<2, 0, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
template __global__ void nuts_kernel<2, 0, 1>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
template __global__ void nuts_kernel<2, 1, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
template __global__ void nuts_kernel<2, 1, 1>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
template __global__ void nuts_kernel<2, 2, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
template __global__ void nuts_kernel<2, 2, 1>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
template __global__ void nuts_kernel<2, 3, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
template __global__ void nuts_kernel<2, 3, 1>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
