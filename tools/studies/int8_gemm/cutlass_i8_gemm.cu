// STUDY (round-2 preparation, not part of libadccb.so): INT8 x INT8 -> INT32 GEMM on the 5th-gen
// tensor cores (tcgen05.mma kind::i8) instantiated from the CUTLASS 4.x sm100 collective builders
// vendored in this image.  C[m,n] = sum_k A[m,k] * B[n,k]  (both operands K-contiguous), exact in
// INT32: the slice product of an Ozaki-type FP64 GEMM emulation (tools/studies/ozaki_int8_*.py).
// Built by tools/studies/int8_gemm/build.py into its own shared object; compiled here, not yet run.
#include <cstdint>

#include "cutlass/cutlass.h"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/util/packed_stride.hpp"
#include "cute/tensor.hpp"

using namespace cute;

using ElementA = int8_t;
using ElementB = int8_t;
using ElementC = int32_t;
using ElementAcc = int32_t;
using LayoutA = cutlass::layout::RowMajor;      // A[m,k], k contiguous
using LayoutB = cutlass::layout::ColumnMajor;   // B[k,n] with k contiguous  ==  B[n,k] row-major
using LayoutC = cutlass::layout::RowMajor;
constexpr int AlignA = 16, AlignB = 16, AlignC = 4;

using TileShape = Shape<_128, _128, _128>;       // M x N x K (K in int8 elements = 128 bytes)
using ClusterShape = Shape<_1, _1, _1>;

using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, TileShape, ClusterShape,
    cutlass::epilogue::collective::EpilogueTileAuto, ElementAcc, ElementAcc, ElementC, LayoutC, AlignC,
    ElementC, LayoutC, AlignC, cutlass::epilogue::collective::EpilogueScheduleAuto>::CollectiveOp;

using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, ElementA, LayoutA, AlignA, ElementB, LayoutB,
    AlignB, ElementAcc, TileShape, ClusterShape,
    cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(
        sizeof(typename CollectiveEpilogue::SharedStorage))>,
    cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;

using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop,
                                                        CollectiveEpilogue, void>;
using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;

extern "C" {

// bytes of device workspace the launch needs
long long study_i8_gemm_workspace(int M, int N, int K) {
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, 1}, {}, {}};
  return (long long)Gemm::get_workspace_size(args);
}

// C (int32, row-major, ld = N) = A (int8 [M,K]) * B^T (int8 [N,K]); returns 0 on success
int study_i8_gemm(void* stream, int M, int N, int K, const int8_t* A, const int8_t* B, int32_t* C,
                  void* workspace) {
  using StrideA = typename Gemm::GemmKernel::StrideA;
  using StrideB = typename Gemm::GemmKernel::StrideB;
  using StrideC = typename Gemm::GemmKernel::StrideC;
  using StrideD = typename Gemm::GemmKernel::StrideD;
  StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, cute::make_shape(M, K, 1));
  StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, cute::make_shape(N, K, 1));
  StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, cute::make_shape(M, N, 1));
  StrideD sd = cutlass::make_cute_packed_stride(StrideD{}, cute::make_shape(M, N, 1));
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm,
                                {M, N, K, 1},
                                {A, sa, B, sb},
                                {{ElementAcc(1), ElementAcc(0)}, C, sc, C, sd}};
  Gemm gemm;
  if (gemm.can_implement(args) != cutlass::Status::kSuccess) return 1;
  if (gemm.initialize(args, workspace, (cudaStream_t)stream) != cutlass::Status::kSuccess) return 2;
  if (gemm.run((cudaStream_t)stream) != cutlass::Status::kSuccess) return 3;
  return 0;
}

}  // extern "C"
