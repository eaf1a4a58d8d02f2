#include "admm_stream.cuh"
namespace mpcb {
int stream_padded(int nt) { return ((nt + 15) / 16) * 16; }
cudaError_t stream_upload(const Design&, StreamConsts&, std::string& err) { err = "streamed kernel not built yet"; return cudaErrorNotSupported; }
cudaError_t stream_solve(const Design&, const mpcb_settings&, const StreamConsts&, StreamWork&, const StreamBatch&, int, cudaStream_t, int*, std::string& err) { err = "streamed kernel not built yet"; return cudaErrorNotSupported; }
void stream_release(StreamConsts&, StreamWork&) {}
}
