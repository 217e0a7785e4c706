// Explicit instantiations of lbfgs_kernel for resident-operand layout TOEP = 1 (see lbfgs_kernel.cuh).
#include "lbfgs_kernel.cuh"

template __global__ void lbfgs_kernel<1, 0, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
template __global__ void lbfgs_kernel<1, 0, 1>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
template __global__ void lbfgs_kernel<1, 1, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
template __global__ void lbfgs_kernel<1, 1, 1>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
template __global__ void lbfgs_kernel<1, 2, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
template __global__ void lbfgs_kernel<1, 2, 1>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
template __global__ void lbfgs_kernel<1, 3, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
template __global__ void lbfgs_kernel<1, 3, 1>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
