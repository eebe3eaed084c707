// sprintz_b200_shim.cpp -- the binding a maintainer of dblalock/sprintz would add.
//
// Defines the eight entry points of the reference's cpp/Compress/sprintz.h:16-32 with the reference's own C++
// signatures (bool write_size = true comes from that header) and forwards each one to the C ABI of
// libsprintz_b200.so (include/sprintz_b200.h).  Compile it INSTEAD of sprintz.cpp + the four codec .cpp files and
// link with -lsprintz_b200: every caller of sprintz.h -- the reference's Catch tests, the lzbench codec table --
// then runs on the GPU path unchanged.  integration/Makefile builds the reference's own test runner
// (cpp/Compress/main.cpp + test/test_sprintz_8b.cpp) this way.
#include "sprintz.h"        // the reference's header, from -I<reference>/cpp/Compress
#include "sprintz_b200.h"   // this repository's C ABI, from -I<repo>/include

int64_t sprintz_compress_delta_8b(const uint8_t* src, uint32_t len, int8_t* dest, uint16_t ndims, bool write_size)
{
    return spz_compress_delta_8b(src, len, dest, ndims, write_size ? 1 : 0);
}
int64_t sprintz_decompress_delta_8b(const int8_t* src, uint8_t* dest) { return spz_decompress_delta_8b(src, dest); }

int64_t sprintz_compress_xff_8b(const uint8_t* src, uint32_t len, int8_t* dest, uint16_t ndims, bool write_size)
{
    return spz_compress_xff_8b(src, len, dest, ndims, write_size ? 1 : 0);
}
int64_t sprintz_decompress_xff_8b(const int8_t* src, uint8_t* dest) { return spz_decompress_xff_8b(src, dest); }

int64_t sprintz_compress_delta_16b(const uint16_t* src, uint32_t len, int16_t* dest, uint16_t ndims, bool write_size)
{
    return spz_compress_delta_16b(src, len, dest, ndims, write_size ? 1 : 0);
}
int64_t sprintz_decompress_delta_16b(const int16_t* src, uint16_t* dest) { return spz_decompress_delta_16b(src, dest); }

int64_t sprintz_compress_xff_16b(const uint16_t* src, uint32_t len, int16_t* dest, uint16_t ndims, bool write_size)
{
    return spz_compress_xff_16b(src, len, dest, ndims, write_size ? 1 : 0);
}
int64_t sprintz_decompress_xff_16b(const int16_t* src, uint16_t* dest) { return spz_decompress_xff_16b(src, dest); }
