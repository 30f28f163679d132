// codecs.cu -- HOST-side block decoders for the .rvdb container.
//
// The reference stores its embeddings in an HDF5 dataset filtered with shuffle + LZF (h5py `compression='lzf',
// shuffle=True`: unified_index.py:943-948, 1557-1562) and its metadata / index blobs as LZ4 frames
// (`lz4.frame.compress`: unified_index.py:950-956, 1827-1860); loading 851 284 x 768 vectors took it 29 s
// (logs/system_20250828.log:26).  Neither h5py nor lz4 exists in this image, so the container is parsed by
// rvdb_reader.py and the bulk bytes are decoded here, in C, one call per chunk / block:
//   * LZF  -- the published liblzf stream format (Marc Lehmann; the h5py filter id 32000 wraps it unchanged);
//   * LZ4  -- the published LZ4 BLOCK format (the frame header / block table is parsed in Python);
//   * byte un-shuffle (HDF5 filter id 2).
// No CUDA in this file; it is compiled into libivr_b200.so with the rest of the C ABI.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../../include/ivr_b200.h"

extern "C" {

// liblzf: ctrl < 32 -> literal run of ctrl + 1 bytes; else back reference: len = ctrl >> 5 (7 -> + next byte), offset =
// ((ctrl & 0x1f) << 8 | next byte) + 1, copy len + 2 bytes.  Returns the decoded size, or -1 on a malformed / oversized stream.
int64_t ivr_lzf_decompress(const uint8_t* src, int64_t src_len, uint8_t* dst, int64_t dst_cap) {
    if (!src || !dst || src_len < 0 || dst_cap < 0) return -1;
    const uint8_t* ip = src;
    const uint8_t* const ie = src + src_len;
    uint8_t* op = dst;
    uint8_t* const oe = dst + dst_cap;
    while (ip < ie) {
        unsigned ctrl = *ip++;
        if (ctrl < 32) {
            const int64_t run = static_cast<int64_t>(ctrl) + 1;
            if (ip + run > ie || op + run > oe) return -1;
            memcpy(op, ip, static_cast<size_t>(run));
            ip += run; op += run;
        } else {
            int64_t len = ctrl >> 5;
            if (len == 7) {
                if (ip >= ie) return -1;
                len += *ip++;
            }
            if (ip >= ie) return -1;
            const int64_t off = (static_cast<int64_t>(ctrl & 0x1f) << 8 | *ip++) + 1;
            len += 2;
            if (op - dst < off || op + len > oe) return -1;
            const uint8_t* ref = op - off;
            for (int64_t i = 0; i < len; ++i) op[i] = ref[i];     // may overlap: byte by byte
            op += len;
        }
    }
    return op - dst;
}

// One LZ4 block, decoded at dst + dst_pos; matches may reach back into dst[0, dst_pos) (linked blocks of a frame).
// Returns the number of bytes produced, or -1 on a malformed / oversized block.
int64_t ivr_lz4_block_decompress(const uint8_t* src, int64_t src_len, uint8_t* dst, int64_t dst_pos, int64_t dst_cap) {
    if (!src || !dst || src_len < 0 || dst_pos < 0 || dst_pos > dst_cap) return -1;
    const uint8_t* ip = src;
    const uint8_t* const ie = src + src_len;
    uint8_t* op = dst + dst_pos;
    uint8_t* const oe = dst + dst_cap;
    while (ip < ie) {
        const unsigned token = *ip++;
        int64_t lit = token >> 4;
        if (lit == 15) {
            unsigned b;
            do { if (ip >= ie) return -1; b = *ip++; lit += b; } while (b == 255);
        }
        if (ip + lit > ie || op + lit > oe) return -1;
        memcpy(op, ip, static_cast<size_t>(lit));
        ip += lit; op += lit;
        if (ip >= ie) break;                                      // the last sequence carries literals only
        if (ip + 2 > ie) return -1;
        const int64_t off = ip[0] | (static_cast<int64_t>(ip[1]) << 8);
        ip += 2;
        int64_t len = token & 15;
        if (len == 15) {
            unsigned b;
            do { if (ip >= ie) return -1; b = *ip++; len += b; } while (b == 255);
        }
        len += 4;
        if (off == 0 || op - dst < off || op + len > oe) return -1;
        const uint8_t* ref = op - off;
        for (int64_t i = 0; i < len; ++i) op[i] = ref[i];
        op += len;
    }
    return op - (dst + dst_pos);
}

// HDF5 shuffle filter, inverse: src holds byte 0 of every element, then byte 1, ...; dst gets the elements back.
int ivr_unshuffle(const uint8_t* src, int64_t n_bytes, int elem_size, uint8_t* dst) {
    if (!src || !dst || n_bytes < 0 || elem_size <= 0) return IVR_EINVAL;
    const int64_t n = n_bytes / elem_size;
    if (elem_size == 4) {
        const uint8_t *b0 = src, *b1 = src + n, *b2 = src + 2 * n, *b3 = src + 3 * n;
        for (int64_t i = 0; i < n; ++i) {
            dst[4 * i] = b0[i]; dst[4 * i + 1] = b1[i]; dst[4 * i + 2] = b2[i]; dst[4 * i + 3] = b3[i];
        }
    } else {
        for (int j = 0; j < elem_size; ++j)
            for (int64_t i = 0; i < n; ++i) dst[i * elem_size + j] = src[j * n + i];
    }
    memcpy(dst + n * elem_size, src + n * elem_size, static_cast<size_t>(n_bytes - n * elem_size));   // leftover bytes are stored as they are
    return IVR_OK;
}

}  // extern "C"
