// Raw DEFLATE decoder for BGZF blocks (inflate_fast.cpp). Product code, host side.
#pragma once
#include <cstddef>
#include <cstdint>

namespace hlac {
// Inflates a raw deflate stream of known inflated size. true: exactly out_len bytes were produced and the final block
// ended. false: anything else (damaged or unusual stream, sizes that do not match) - out may be partially written; the
// caller falls back to zlib and checks the block's CRC32 either way. Never reads outside [in, in + in_len) or writes outside
// [out, out + out_len).
bool inflate_raw_fast(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);
}  // namespace hlac
