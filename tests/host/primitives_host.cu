// Host build (nvcc, no GPU needed) of the __host__ __device__ primitives in sprintz_b200/csrc/spz_common.cuh, so the CPU
// test suite can check them exhaustively the way the reference's test_bitpack.cpp does for its own (zigzag over all 256 /
// 65536 values).  Test infrastructure, not product.
#include "../../sprintz_b200/csrc/spz_common.cuh"
#include "../../sprintz_b200/csrc/spz_swar.cuh"

extern "C" {
uint32_t h_zigzag_enc(int w, int32_t e) { return w == 8 ? spz::zigzag_enc<8>(e) : spz::zigzag_enc<16>(e); }
int32_t h_zigzag_dec(int w, uint32_t u) { return w == 8 ? spz::zigzag_dec<8>(u) : spz::zigzag_dec<16>(u); }
int32_t h_fire_coef(int w, int low, int32_t c)
{
    if (w == 8) return low ? spz::fire_coef<8, true>(c) : spz::fire_coef<8, false>(c);
    return low ? spz::fire_coef<16, true>(c) : spz::fire_coef<16, false>(c);
}
int32_t h_fire_predict(int w, int32_t pd, int32_t coef) { return w == 8 ? spz::fire_predict<8>(pd, coef) : spz::fire_predict<16>(pd, coef); }
int32_t h_fire_grad(int w, int32_t err, int32_t pd) { return w == 8 ? spz::fire_grad<8>(err, pd) : spz::fire_grad<16>(err, pd); }
// op: 0 psub, 1 padd, 2 pzigzag (b unused), 3 punzigzag (b unused); n words
void h_swar(int w, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, uint64_t n)
{
    for (uint64_t i = 0; i < n; i++) {
        uint32_t r;
        if (w == 8) r = op == 0 ? spz::psub<8>(a[i], b[i]) : op == 1 ? spz::padd<8>(a[i], b[i]) : op == 2 ? spz::pzigzag<8>(a[i]) : spz::punzigzag<8>(a[i]);
        else        r = op == 0 ? spz::psub<16>(a[i], b[i]) : op == 1 ? spz::padd<16>(a[i], b[i]) : op == 2 ? spz::pzigzag<16>(a[i]) : spz::punzigzag<16>(a[i]);
        out[i] = r;
    }
}
int32_t h_sx(int bits, int32_t v) { return bits == 8 ? spz::sx<8>(v) : bits == 16 ? spz::sx<16>(v) : spz::sx<32>(v); }
}
