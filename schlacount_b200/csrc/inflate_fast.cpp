// Raw DEFLATE (RFC 1951) decoder for BGZF blocks, written for the host BAM reader (bam_reader.cpp): a BAM file is a chain
// of <= 64 KB deflate streams and inflating them is more than half of the host work per record. Every block
// carries its inflated size and a CRC32, so the decoder is allowed to be optimistic: it either produces exactly out_len
// bytes and returns true, or returns false - the caller then runs zlib's inflate on the block, and checks the CRC32 in both
// cases. Product code (host side); no reference counterpart (the reference goes through rust-htslib).
//
// Design: 64-bit bit buffer refilled with one unaligned 8-byte load per literal/length + distance pair (48 bits at most);
// two-level decode tables (11 bits for literals/lengths, 8 for distances, sub-tables for longer codes) whose entries carry
// everything the loop needs (bits to drop, extra-bit count, base value); a fast loop without bounds checks while at least
// 8 input bytes and 258 + 40 output bytes remain, a checked loop for the ends; matches copied 8 bytes at a time when the
// distance allows.
#include "inflate_fast.hpp"

#include <cstring>

namespace hlac {
namespace {

constexpr int LIT_BITS = 11, DIST_BITS = 8, PRE_BITS = 7;
constexpr int LIT_ENOUGH = 2048 + 1024, DIST_ENOUGH = 256 + 512;   // main table + the most sub-table entries 15-bit codes can need

// entry: bits 0..7 = bits to drop (code length; for a sub-table pointer the main-table bits), 8..15 = kind / extra bits,
// 16..31 = literal, base value or sub-table start
constexpr uint32_t K_LITERAL = 0x80, K_EOB = 0x40, K_SUB = 0x20, K_EXTRA_MASK = 0x1F;
inline uint32_t mk(uint32_t value, uint32_t kind, uint32_t bits) { return (value << 16) | (kind << 8) | bits; }

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

// the low `len` (<= 15) bits of code in reverse order
struct Rev8 { uint8_t v[256]; constexpr Rev8() : v{} { for (int b = 0; b < 256; b++) { int r = 0; for (int i = 0; i < 8; i++) r |= ((b >> i) & 1) << (7 - i); v[b] = (uint8_t)r; } } };
constexpr Rev8 REV8{};
inline uint32_t bitrev(uint32_t code, int len) {
  const uint32_t r16 = ((uint32_t)REV8.v[code & 0xFF] << 8) | REV8.v[(code >> 8) & 0xFF];
  return r16 >> (16 - len);
}

enum TableKind { T_PRE, T_LITLEN, T_DIST };

// what a symbol decodes to, as a table entry (without the bits-to-drop field). Symbols that take part in a code but must
// never occur (286, 287, distance 30, 31 of the fixed code) become error entries: K_EOB with a non-zero value.
constexpr uint32_t ERR_VALUE = 1;
inline void symbol_entry(TableKind kind, uint32_t sym, uint32_t& value, uint32_t& k) {
  if (kind == T_PRE) { value = sym; k = 0; return; }
  if (kind == T_LITLEN) {
    if (sym < 256) { value = sym; k = K_LITERAL; }
    else if (sym == 256) { value = 0; k = K_EOB; }
    else if (sym <= 285) { value = LEN_BASE[sym - 257]; k = LEN_EXTRA[sym - 257]; }
    else { value = ERR_VALUE; k = K_EOB; }
    return;
  }
  if (sym <= 29) { value = DIST_BASE[sym]; k = DIST_EXTRA[sym]; }
  else { value = ERR_VALUE; k = K_EOB; }
}

// Canonical Huffman code -> two-level table. false for an over-subscribed code, an incomplete code (except the one case
// RFC 1951 allows in practice: a single distance code of length 1, and the all-zero distance code of a literal-only block)
// or a table that would not fit.
bool build_table(TableKind kind, const uint8_t* lens, int n, int table_bits, uint32_t* table, int capacity) {
  uint16_t count[16] = {0};
  for (int i = 0; i < n; i++) count[lens[i]]++;
  if (count[0] == n) {   // no codes at all
    if (kind != T_DIST) return false;
    for (int i = 0; i < (1 << table_bits); i++) table[i] = mk(ERR_VALUE, K_EOB, 0);   // any use of a distance is an error
    return true;
  }
  int left = 1;
  for (int len = 1; len <= 15; len++) { left <<= 1; left -= count[len]; if (left < 0) return false; }
  bool incomplete = left > 0;
  if (incomplete && !(kind == T_DIST && count[1] == 1 && n - count[0] == 1)) return false;
  // symbols sorted by (length, symbol)
  uint16_t offs[16]; offs[1] = 0;
  for (int len = 1; len < 15; len++) offs[len + 1] = (uint16_t)(offs[len] + count[len]);
  uint16_t sorted[288];
  if (n > 288) return false;
  for (int i = 0; i < n; i++) if (lens[i]) sorted[offs[lens[i]]++] = (uint16_t)i;
  const int n_codes = n - count[0];
  const int main_size = 1 << table_bits;
  if (incomplete) for (int i = 0; i < main_size; i++) table[i] = mk(ERR_VALUE, K_EOB, 0);   // the unused half of a one-code distance table: an error when hit
  int next_free = main_size;
  uint32_t code = 0;
  int len = 1;
  while (count[len] == 0) { len++; code <<= 1; }
  int remaining_in_len = count[len];
  for (int k = 0; k < n_codes;) {
    const uint32_t sym = sorted[k];
    uint32_t value, kk;
    symbol_entry(kind, sym, value, kk);
    if (len <= table_bits) {
      const uint32_t rev = bitrev(code, len);
      const uint32_t e = mk(value, kk, (uint32_t)len);
      for (uint32_t i = rev; i < (uint32_t)main_size; i += 1u << len) table[i] = e;
      k++;
    } else {
      // all codes sharing the first table_bits bits form a run in canonical order; its last code is its longest
      const uint32_t prefix = code >> (len - table_bits);
      int k2 = k, len2 = len, rem2 = remaining_in_len; uint32_t code2 = code; int max_len = len;
      while (k2 < n_codes && (code2 >> (len2 - table_bits)) == prefix) {
        max_len = len2;
        k2++; code2++;
        if (--rem2 == 0) { do { len2++; code2 <<= 1; } while (len2 <= 15 && count[len2] == 0); if (len2 > 15) break; rem2 = count[len2]; }
      }
      const int sub_bits = max_len - table_bits;
      const int sub_size = 1 << sub_bits;
      if (next_free + sub_size > capacity) return false;
      const uint32_t rev_prefix = bitrev(prefix, table_bits);
      table[rev_prefix] = mk((uint32_t)next_free, K_SUB | (uint32_t)sub_bits, (uint32_t)table_bits);
      // an incomplete sub-table cannot happen for a complete code, but fill it so that a damaged stream cannot read garbage
      for (int i = 0; i < sub_size; i++) table[next_free + i] = mk(ERR_VALUE, K_EOB, 0);
      while (k < k2) {
        const uint32_t s2 = sorted[k];
        symbol_entry(kind, s2, value, kk);
        const int sl = len - table_bits;                       // bits of this code inside the sub-table
        const uint32_t rev = bitrev(code & ((1u << sl) - 1), sl);
        const uint32_t e = mk(value, kk, (uint32_t)sl);
        for (uint32_t i = rev; i < (uint32_t)sub_size; i += 1u << sl) table[next_free + i] = e;
        k++; code++;
        if (--remaining_in_len == 0) { do { len++; code <<= 1; } while (len <= 15 && count[len] == 0); if (len <= 15) remaining_in_len = count[len]; }
      }
      next_free += sub_size;
      continue;
    }
    code++;
    if (--remaining_in_len == 0) { do { len++; code <<= 1; } while (len <= 15 && count[len] == 0); if (len <= 15) remaining_in_len = count[len]; }
  }
  return true;
}

inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }   // little-endian hosts only (x86-64 / aarch64 LE)

struct Tables { uint32_t lit[LIT_ENOUGH]; uint32_t dist[DIST_ENOUGH]; };

}  // namespace

bool inflate_raw_fast(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
  const uint8_t* const in_end = in + in_len;
  uint8_t* const out_begin = out;
  uint8_t* const out_end = out + out_len;
  uint64_t bitbuf = 0;
  unsigned bitcnt = 0;
  Tables t;
  static const struct Fixed {   // the fixed code of BTYPE 01, built once
    Tables t; bool ok;
    Fixed() {
      uint8_t l[288], d[32];
      for (int i = 0; i < 144; i++) l[i] = 8;
      for (int i = 144; i < 256; i++) l[i] = 9;
      for (int i = 256; i < 280; i++) l[i] = 7;
      for (int i = 280; i < 288; i++) l[i] = 8;
      for (int i = 0; i < 32; i++) d[i] = 5;
      ok = build_table(T_LITLEN, l, 288, LIT_BITS, t.lit, LIT_ENOUGH) && build_table(T_DIST, d, 32, DIST_BITS, t.dist, DIST_ENOUGH);
    }
  } fixed;

#define HLAC_REFILL()                                                                                          \
  do {                                                                                                         \
    if (in_end - in >= 8) { bitbuf |= load64(in) << bitcnt; in += (63 - bitcnt) >> 3; bitcnt |= 56; }           \
    else while (bitcnt <= 56 && in < in_end) { bitbuf |= (uint64_t)*in++ << bitcnt; bitcnt += 8; }              \
  } while (0)
#define HLAC_NEED(n) do { if (bitcnt < (unsigned)(n)) return false; } while (0)
#define HLAC_DROP(n) do { bitbuf >>= (n); bitcnt -= (n); } while (0)

  for (;;) {
    HLAC_REFILL();
    HLAC_NEED(3);
    const unsigned bfinal = (unsigned)bitbuf & 1, btype = ((unsigned)bitbuf >> 1) & 3;
    HLAC_DROP(3);
    const uint32_t* lit; const uint32_t* dist;
    if (btype == 0) {   // stored: back to the byte boundary, LEN / NLEN, raw bytes
      HLAC_DROP(bitcnt & 7);
      in -= bitcnt >> 3;   // whole bytes still in the bit buffer go back to the input
      bitbuf = 0; bitcnt = 0;
      if (in_end - in < 4) return false;
      const unsigned len = in[0] | (in[1] << 8), nlen = in[2] | (in[3] << 8);
      in += 4;
      if ((len ^ nlen) != 0xFFFF) return false;
      if ((size_t)(in_end - in) < len || (size_t)(out_end - out) < len) return false;
      memcpy(out, in, len);
      in += len; out += len;
      if (bfinal) break;
      continue;
    } else if (btype == 1) {
      if (!fixed.ok) return false;
      lit = fixed.t.lit; dist = fixed.t.dist;
    } else if (btype == 2) {
      HLAC_NEED(14);
      const unsigned hlit = ((unsigned)bitbuf & 31) + 257, hdist = (((unsigned)bitbuf >> 5) & 31) + 1, hclen = (((unsigned)bitbuf >> 10) & 15) + 4;
      HLAC_DROP(14);
      if (hlit > 286 || hdist > 30) return false;
      static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint8_t pre_lens[19] = {0};
      for (unsigned i = 0; i < hclen; i++) {
        if (bitcnt < 3) HLAC_REFILL();
        HLAC_NEED(3);
        pre_lens[order[i]] = (uint8_t)(bitbuf & 7);
        HLAC_DROP(3);
      }
      uint32_t pre[1 << PRE_BITS];
      if (!build_table(T_PRE, pre_lens, 19, PRE_BITS, pre, 1 << PRE_BITS)) return false;
      uint8_t lens[286 + 30 + 138];
      unsigned i = 0;
      while (i < hlit + hdist) {
        if (bitcnt < 7 + 7) HLAC_REFILL();
        const uint32_t e = pre[bitbuf & ((1u << PRE_BITS) - 1)];
        const unsigned nb = e & 0xFF;
        if (nb > bitcnt || (e >> 8 & K_EOB)) return false;
        HLAC_DROP(nb);
        const unsigned sym = e >> 16;
        if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
        unsigned rep, val = 0;
        if (sym == 16) { if (i == 0) return false; HLAC_NEED(2); val = lens[i - 1]; rep = 3 + ((unsigned)bitbuf & 3); HLAC_DROP(2); }
        else if (sym == 17) { HLAC_NEED(3); rep = 3 + ((unsigned)bitbuf & 7); HLAC_DROP(3); }
        else { HLAC_NEED(7); rep = 11 + ((unsigned)bitbuf & 127); HLAC_DROP(7); }
        if (i + rep > hlit + hdist) return false;
        memset(lens + i, (int)val, rep);
        i += rep;
      }
      if (lens[256] == 0) return false;   // no end-of-block code
      if (!build_table(T_LITLEN, lens, (int)hlit, LIT_BITS, t.lit, LIT_ENOUGH)) return false;
      if (!build_table(T_DIST, lens + hlit, (int)hdist, DIST_BITS, t.dist, DIST_ENOUGH)) return false;
      lit = t.lit; dist = t.dist;
    } else return false;

    // ---- symbols of one block ----
    for (;;) {
      // fast loop: no bounds checks while both ends are far away (one refill covers a length + distance pair: <= 48 bits)
      while (in_end - in >= 8 && out_end - out >= 258 + 40) {
        bitbuf |= load64(in) << bitcnt; in += (63 - bitcnt) >> 3; bitcnt |= 56;
        uint32_t e = lit[bitbuf & ((1u << LIT_BITS) - 1)];
        if (e & (K_SUB << 8)) { HLAC_DROP(LIT_BITS); e = lit[(e >> 16) + (bitbuf & ((1u << ((e >> 8) & K_EXTRA_MASK)) - 1))]; }
        HLAC_DROP(e & 0xFF);
        if (e & (K_LITERAL << 8)) {
          *out++ = (uint8_t)(e >> 16);
          // a second and third literal from the same refill (<= 3 x 15 bits)
          e = lit[bitbuf & ((1u << LIT_BITS) - 1)];
          if (!(e & (K_LITERAL << 8))) continue;   // decoded again at the top (nothing was dropped)
          HLAC_DROP(e & 0xFF);
          *out++ = (uint8_t)(e >> 16);
          e = lit[bitbuf & ((1u << LIT_BITS) - 1)];
          if (!(e & (K_LITERAL << 8))) continue;
          HLAC_DROP(e & 0xFF);
          *out++ = (uint8_t)(e >> 16);
          continue;
        }
        if (e & (K_EOB << 8)) { if ((e >> 16) != 0) return false; goto block_done; }
        const unsigned xb = (e >> 8) & K_EXTRA_MASK;
        const unsigned length = (e >> 16) + ((unsigned)bitbuf & ((1u << xb) - 1));
        HLAC_DROP(xb);
        uint32_t d = dist[bitbuf & ((1u << DIST_BITS) - 1)];
        if (d & (K_SUB << 8)) { HLAC_DROP(DIST_BITS); d = dist[(d >> 16) + (bitbuf & ((1u << ((d >> 8) & K_EXTRA_MASK)) - 1))]; }
        if (d & (K_EOB << 8)) return false;   // unused distance code
        HLAC_DROP(d & 0xFF);
        const unsigned db = (d >> 8) & K_EXTRA_MASK;
        const size_t distance = (d >> 16) + ((size_t)bitbuf & (((size_t)1 << db) - 1));
        HLAC_DROP(db);
        if (distance > (size_t)(out - out_begin)) return false;
        const uint8_t* src = out - distance;
        uint8_t* dst = out;
        out += length;
        if (distance >= 16) {   // 16 bytes at a time; may write up to 15 bytes past the match (room was checked)
          do { memcpy(dst, src, 16); dst += 16; src += 16; } while (dst < out);
        } else if (distance >= 8) {
          do { memcpy(dst, src, 8); dst += 8; src += 8; } while (dst < out);
        } else if (distance == 1) {
          memset(dst, *src, length);
        } else {
          do { *dst++ = *src++; } while (dst < out);
        }
      }
      // checked step for the ends of the block / stream
      HLAC_REFILL();
      uint32_t e = lit[bitbuf & ((1u << LIT_BITS) - 1)];
      if (e & (K_SUB << 8)) { HLAC_NEED(LIT_BITS); HLAC_DROP(LIT_BITS); e = lit[(e >> 16) + (bitbuf & ((1u << ((e >> 8) & K_EXTRA_MASK)) - 1))]; }
      if ((e & 0xFF) > bitcnt) return false;
      HLAC_DROP(e & 0xFF);
      if (e & (K_LITERAL << 8)) { if (out == out_end) return false; *out++ = (uint8_t)(e >> 16); continue; }
      if (e & (K_EOB << 8)) { if ((e >> 16) != 0) return false; goto block_done; }
      const unsigned xb = (e >> 8) & K_EXTRA_MASK;
      HLAC_NEED(xb);
      const unsigned length = (e >> 16) + ((unsigned)bitbuf & ((1u << xb) - 1));
      HLAC_DROP(xb);
      uint32_t d = dist[bitbuf & ((1u << DIST_BITS) - 1)];
      if (d & (K_SUB << 8)) { HLAC_NEED(DIST_BITS); HLAC_DROP(DIST_BITS); d = dist[(d >> 16) + (bitbuf & ((1u << ((d >> 8) & K_EXTRA_MASK)) - 1))]; }
      if (d & (K_EOB << 8)) return false;
      if ((d & 0xFF) > bitcnt) return false;
      HLAC_DROP(d & 0xFF);
      const unsigned db = (d >> 8) & K_EXTRA_MASK;
      HLAC_NEED(db);
      const size_t distance = (d >> 16) + ((size_t)bitbuf & (((size_t)1 << db) - 1));
      HLAC_DROP(db);
      if (distance > (size_t)(out - out_begin) || length > (size_t)(out_end - out)) return false;
      const uint8_t* src = out - distance;
      for (unsigned k = 0; k < length; k++) out[k] = src[k];
      out += length;
    }
  block_done:
    if (bfinal) break;
  }
#undef HLAC_REFILL
#undef HLAC_NEED
#undef HLAC_DROP
  return out == out_end;
}

}  // namespace hlac
