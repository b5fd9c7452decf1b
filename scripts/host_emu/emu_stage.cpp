// DEVELOPER ANALYSIS AID: how many passes through `pow` would a warp need if its lanes ran the
// routing of a segment as a state machine (one pass = every lane's next power, whatever its stage)?
#define HELP_EMU_STAGE
#define EMU_MAXSEG 16
#define EMU_REGS 80
#include "emu.cpp"
extern "C" long long stage_len(int cell) { return (long long)g_stage[cell].size(); }
extern "C" void stage_get(int cell, unsigned char* out) { auto& v = g_stage[cell]; for (size_t i = 0; i < v.size(); ++i) out[i] = v[i]; }
extern "C" void stage_keys(const help_b200_batch* b, unsigned long long* key) {
  const int64_t n = b->n_cells; const int ny = b->n_years;
  BatchDev B{};
  B.n_cells = n; B.n_years = ny; B.max_layers = b->max_layers; B.n_nodes_p = b->n_nodes_p; B.n_nodes_t = b->n_nodes_t; B.n_nodes_r = b->n_nodes_r;
  B.years = b->years; B.pre = (float*)b->pre; B.tmp = (float*)b->tmp; B.rad = (float*)b->rad; B.iu4 = b->iu4; B.iu7 = b->iu7; B.iu13 = b->iu13;
  B.tm_normals = b->tm_normals; B.idfs = b->idfs; B.cell_f32 = b->cell_f32; B.cell_i32 = b->cell_i32; B.layer_f32 = b->layer_f32; B.layer_i32 = b->layer_i32; B.layer_rc = b->layer_rc;
  int segcap = 7 + 3 * b->max_layers; if (segcap > 67) segcap = 67;
  int64_t stride = ((n + 31) / 32) * 32;
  std::vector<float> f32(Image::n_f32(segcap) * stride); std::vector<int32_t> i32(Image::n_i32(segcap) * stride); std::vector<double> f64(Image::n_f64(segcap) * stride);
  std::vector<char> rec((size_t)Image::rec_bytes(stride, segcap));
  Image img{f32.data(), i32.data(), f64.data(), stride, segcap, rec.data()};
  blockDim.x = 1; gridDim.x = 1; threadIdx.x = 0;
  for (int64_t c = 0; c < n; ++c) { blockIdx.x = (unsigned)c; help_setup_kernel(B, img, nullptr, nullptr, nullptr, (uint64_t*)key); }
}
