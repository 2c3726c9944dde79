// Host interface of the tensor-core DFT frontend kernel (frontend_tc.cu), used by frontend.cu's launch path.
#pragma once

#include <cuda_runtime.h>

#include "frontend_params.cuh"
#include "sir_common.cuh"

namespace sir {

// Builds the operand images / twiddles / unscaled mel taps on the host and uploads them into `buf`.
int frontend_tc_upload_tables(DeviceBuffer& buf, TcDeviceTables& dev, int sample_rate, int n_mels);
// Work items (15-frame tiles) of an utterance of n_frames frames, and the tickets one launch consumes.
int frontend_tc_groups(int n_frames);
long long frontend_tc_tickets(long long items, long long grid);
// Launches the kernel (one persistent CTA per SM, grid = min(items, num_sms)).  The caller checks cudaGetLastError.
int frontend_tc_launch(const FrontendParams& p, bool pcm16, int num_sms, cudaStream_t stream);

}  // namespace sir
