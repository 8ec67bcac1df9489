// JPEG reader for the reference's image inputs (Textures::Textures, src/Base/src/Textures.cpp:27-58: the earth albedo, star map, moon albedo and
// moon normal maps under data/NASA are loaded with stbi_load; all four are PROGRESSIVE 4:4:4 YCbCr files, two of them with successive
// approximation).  stb_image (external/stb/stb_image.h) is a vendored third-party header of the reference; this restates the published format it
// implements for this path from the specification -- ITU-T T.81: baseline and progressive DCT, Huffman coding, restart intervals, 8-bit
// samples, 1 or 3 components with sampling factors 1 or 2 -- and reproduces stb_image's ARITHMETIC where T.81 leaves the decoder free, so that a
// texture loaded here holds the same bytes the reference uploads:
//   * the inverse DCT: the Loeffler-Ligtenberg-Moschytz factorisation with 12-bit constants, rounding 512 >> 10 after the column pass and
//     (65536 + (128 << 17)) >> 17 after the row pass (stb_image.h, stbi__idct_block);
//   * chroma upsampling: the 3:1 triangle filters of the h2 / v2 / h2v2 cases (stbi__resample_row_*);
//   * YCbCr -> RGB in 20-bit fixed point with the blue-difference term of green masked to its upper 16 bits (stbi__YCbCr_to_RGB_row).
// Pinned: where the reference tree is mounted, tests/test_jpeg.py compiles that header into oracle/_ref/libstbref.so and compares every NASA map and a
// matrix of generated files (baseline / progressive x 4:4:4 / 4:2:2 / 4:2:0 / grey x restart intervals) byte for byte; digests of the four maps are committed.
// Not supported (fails with a message): arithmetic coding, lossless / hierarchical modes, 12-bit samples, CMYK / YCCK, sampling factors above 2.
#include "jpeg.h"

#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace skyhost {
namespace {

[[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("jpeg: ") + what); }

// T.81 figure A.6: position k of the zig-zag sequence -> index into the 8x8 block in natural (row-major) order
constexpr uint8_t kNatural[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                  35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// T.81 annex C / F.2.2.3: canonical Huffman code, decoded through per-length (min code, max code, first symbol) tables
struct HuffmanTable {
    bool defined = false;
    uint8_t symbols[256] = {};
    int32_t max_code[18] = {};   // largest code of each length, left-aligned to 16 bits; -1: no code of that length
    int32_t first_code[17] = {}, first_symbol[17] = {};
    void build(const uint8_t counts[16], const uint8_t* values, int total) {
        std::memcpy(symbols, values, size_t(total));
        int32_t code = 0, k = 0;
        for (int len = 1; len <= 16; ++len) {
            first_symbol[len] = k;
            first_code[len] = code;
            if (counts[len - 1]) {
                if (code + counts[len - 1] - 1 >= (1 << len)) fail("bad Huffman code lengths");
                k += counts[len - 1];
                code += counts[len - 1];
                max_code[len] = ((code - 1) << (16 - len)) | ((1 << (16 - len)) - 1);
            } else {
                max_code[len] = -1;
            }
            code <<= 1;
        }
        max_code[17] = 0x7fffffff;
        defined = true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0;       // SOF: identifier, sampling factors, quantisation table
    int td = 0, ta = 0;                     // SOS: DC / AC Huffman table
    int dc_pred = 0;
    int x = 0, y = 0;                       // samples of this component inside the image
    int blocks_w = 0, blocks_h = 0;         // blocks allocated (whole MCUs)
    std::vector<int16_t> coef;              // progressive: [blocks_h][blocks_w][64], natural order, before dequantisation
    std::vector<uint8_t> plane;             // decoded samples, stride blocks_w * 8
};

struct Decoder {
    const uint8_t* p;
    size_t n, pos = 0;
    // entropy-coded segment reader (T.81 F.2.2.5): MSB first, 0xFF00 -> 0xFF, a marker ends the data (zeros are fed from there on)
    uint32_t acc = 0;
    int bits = 0;
    int pending_marker = -1;

    int width = 0, height = 0;
    bool progressive = false;
    int ncomp = 0, h_max = 1, v_max = 1, mcus_x = 0, mcus_y = 0;
    Component comp[3];
    uint16_t quant[4][64] = {};   // natural order
    bool quant_defined[4] = {};
    HuffmanTable dc_table[4], ac_table[4];
    int restart_interval = 0;
    int adobe_transform = -1;     // APP14: 0 = the components are RGB (or CMYK), 1 = YCbCr
    bool jfif = false;
    // current scan
    int scan_n = 0, scan_comp[3] = {}, ss = 0, se = 63, ah = 0, al = 0, eob_run = 0;

    uint8_t byte() { if (pos >= n) fail("file is truncated"); return p[pos++]; }
    int be16() { int hi = byte(); return (hi << 8) | byte(); }

    void fill() {
        while (bits <= 24) {
            uint32_t b = 0;
            if (pending_marker < 0 && pos < n) {
                b = p[pos++];
                if (b == 0xff) {
                    uint8_t c = pos < n ? p[pos++] : 0xd9;
                    while (c == 0xff && pos < n) c = p[pos++];   // fill bytes before a marker
                    if (c != 0) { pending_marker = c; b = 0; }
                }
            }
            acc |= b << (24 - bits);
            bits += 8;
        }
    }
    int take(int k) {   // k <= 16
        if (k == 0) return 0;
        if (bits < k) fill();
        int v = int(acc >> (32 - k));
        acc <<= k; bits -= k;
        return v;
    }
    int bit() { return take(1); }
    int huffman(const HuffmanTable& t) {
        if (!t.defined) fail("scan uses an undefined Huffman table");
        if (bits < 16) fill();
        const int32_t top = int32_t(acc >> 16);
        int len = 1;
        while (top > t.max_code[len]) ++len;
        if (len > 16) fail("bad Huffman code");
        const int code = int(acc >> (32 - len));
        acc <<= len; bits -= len;
        return t.symbols[t.first_symbol[len] + code - t.first_code[len]];
    }
    // T.81 F.2.2.1 (EXTEND of the next s bits)
    int receive_extend(int s) {
        if (s == 0) return 0;
        const int v = take(s);
        return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
    }
    void reset_entropy() {
        acc = 0; bits = 0; pending_marker = -1; eob_run = 0;
        for (int c = 0; c < ncomp; ++c) comp[c].dc_pred = 0;
    }

    // ---- marker segments (T.81 annex B) ------------------------------------------------------------------------------------
    void read_dqt(int len) {
        while (len > 0) {
            const int pq_tq = byte(), precision = pq_tq >> 4, t = pq_tq & 15;
            if (precision > 1 || t > 3) fail("bad quantisation table");
            for (int k = 0; k < 64; ++k) quant[t][kNatural[k]] = uint16_t(precision ? be16() : byte());
            quant_defined[t] = true;
            len -= precision ? 129 : 65;
        }
        if (len != 0) fail("bad DQT segment");
    }
    void read_dht(int len) {
        while (len > 0) {
            const int tc_th = byte(), cls = tc_th >> 4, t = tc_th & 15;
            if (cls > 1 || t > 3) fail("bad Huffman table header");
            uint8_t counts[16], values[256];
            int total = 0;
            for (int i = 0; i < 16; ++i) total += counts[i] = byte();
            if (total > 256) fail("bad Huffman table");
            for (int i = 0; i < total; ++i) values[i] = byte();
            (cls ? ac_table : dc_table)[t].build(counts, values, total);
            len -= 17 + total;
        }
        if (len != 0) fail("bad DHT segment");
    }
    void read_sof(int marker, int len) {
        if (ncomp) fail("more than one frame");
        if (marker != 0xc0 && marker != 0xc1 && marker != 0xc2) fail("only baseline, extended sequential and progressive Huffman DCT frames are supported");
        progressive = marker == 0xc2;
        if (byte() != 8) fail("only 8-bit samples are supported");
        height = be16(); width = be16();
        if (height == 0) fail("a frame without height (DNL) is not supported");
        if (width == 0) fail("zero width");
        ncomp = byte();
        if (ncomp != 1 && ncomp != 3) { ncomp = 0; fail("only 1- and 3-component images are supported"); }
        if (len != 8 + 3 * ncomp) fail("bad SOF segment");
        for (int c = 0; c < ncomp; ++c) {
            Component& k = comp[c];
            k.id = byte();
            const int hv = byte();
            k.h = hv >> 4; k.v = hv & 15; k.tq = byte();
            if (k.h < 1 || k.h > 2 || k.v < 1 || k.v > 2) fail("sampling factors other than 1 and 2 are not supported");
            if (k.tq > 3) fail("bad quantisation table index");
            h_max = std::max(h_max, k.h); v_max = std::max(v_max, k.v);
        }
        if (int64_t(width) * height * ncomp > (int64_t(1) << 31)) fail("image too large");
        mcus_x = (width + 8 * h_max - 1) / (8 * h_max);
        mcus_y = (height + 8 * v_max - 1) / (8 * v_max);
        for (int c = 0; c < ncomp; ++c) {
            Component& k = comp[c];
            k.x = (width * k.h + h_max - 1) / h_max;
            k.y = (height * k.v + v_max - 1) / v_max;
            k.blocks_w = mcus_x * k.h; k.blocks_h = mcus_y * k.v;
            k.plane.assign(size_t(k.blocks_w) * 8 * k.blocks_h * 8, 0);
            if (progressive) k.coef.assign(size_t(k.blocks_w) * k.blocks_h * 64, 0);
        }
    }
    void read_sos(int len) {
        if (!ncomp) fail("scan before the frame header");
        scan_n = byte();
        if (scan_n < 1 || scan_n > ncomp || len != 6 + 2 * scan_n) fail("bad SOS segment");
        for (int i = 0; i < scan_n; ++i) {
            const int id = byte(), tables = byte();
            int c = 0;
            while (c < ncomp && comp[c].id != id) ++c;
            if (c == ncomp) fail("scan names an unknown component");
            comp[c].td = tables >> 4; comp[c].ta = tables & 15;
            if (comp[c].td > 3 || comp[c].ta > 3) fail("bad Huffman table index");
            scan_comp[i] = c;
        }
        ss = byte(); se = byte();
        const int a = byte();
        ah = a >> 4; al = a & 15;
        if (progressive) {
            if (ss > 63 || se > 63 || ss > se || ah > 13 || al > 13) fail("bad progressive scan parameters");
            if (ss == 0 && se != 0) fail("a progressive scan mixes DC and AC coefficients");
            if (ss > 0 && scan_n != 1) fail("a progressive AC scan must have one component");
        } else {
            if (ss != 0 || ah != 0 || al != 0) fail("bad sequential scan parameters");
            se = 63;
        }
    }

    // ---- block decoding -------------------------------------------------------------------------------------------------------
    // sequential (T.81 F.2.2): coefficients dequantised as they are decoded, natural order
    void block_sequential(Component& k, int16_t* out) {
        std::memset(out, 0, 64 * sizeof(int16_t));
        if (!quant_defined[k.tq]) fail("frame uses an undefined quantisation table");
        const uint16_t* q = quant[k.tq];
        const int t = huffman(dc_table[k.td]);
        if (t > 15) fail("bad DC code");
        k.dc_pred += receive_extend(t);
        out[0] = int16_t(k.dc_pred * q[0]);
        const HuffmanTable& ac = ac_table[k.ta];
        for (int i = 1; i < 64;) {
            const int rs = huffman(ac), run = rs >> 4, size = rs & 15;
            if (size == 0) {
                if (rs != 0xf0) break;   // EOB
                i += 16;
            } else {
                i += run;
                if (i > 63) fail("coefficient index out of range");
                const int pos_n = kNatural[i++];
                out[pos_n] = int16_t(receive_extend(size) * q[pos_n]);
            }
        }
    }
    // progressive DC (T.81 G.1.2.1): first scan sets the coefficient to the prediction << Al, a refinement adds bit Al
    void block_progressive_dc(Component& k, int16_t* c) {
        if (ah == 0) {
            std::memset(c, 0, 64 * sizeof(int16_t));
            const int t = huffman(dc_table[k.td]);
            if (t > 15) fail("bad DC code");
            k.dc_pred += receive_extend(t);
            c[0] = int16_t(k.dc_pred * (1 << al));
        } else if (bit()) {
            c[0] = int16_t(c[0] + (1 << al));
        }
    }
    // progressive AC (T.81 G.1.2.2, G.1.2.3) of the band [ss, se]
    void block_progressive_ac(Component& k, int16_t* c) {
        const HuffmanTable& ac = ac_table[k.ta];
        if (ah == 0) {
            if (eob_run) { --eob_run; return; }
            for (int i = ss; i <= se;) {
                const int rs = huffman(ac), run = rs >> 4, size = rs & 15;
                if (size == 0) {
                    if (run < 15) {   // EOBn: this block and 2^run + extra - 1 more end here
                        eob_run = (1 << run) - 1;
                        if (run) eob_run += take(run);
                        break;
                    }
                    i += 16;
                } else {
                    i += run;
                    if (i > 63) fail("coefficient index out of range");
                    c[kNatural[i++]] = int16_t(receive_extend(size) * (1 << al));
                }
            }
            return;
        }
        // refinement: every coefficient with history gets a correction bit; new coefficients are +-(1 << al)
        const int16_t delta = int16_t(1 << al);
        auto refine = [&](int16_t& v) {
            if (bit() && (v & delta) == 0) v = int16_t(v > 0 ? v + delta : v - delta);
        };
        int i = ss;
        if (eob_run == 0) {
            while (i <= se) {
                const int rs = huffman(ac), size = rs & 15;
                int run = rs >> 4, value = 0;
                if (size == 0) {
                    if (run < 15) {
                        eob_run = (1 << run);
                        if (run) eob_run += take(run);
                        break;   // the rest of this block is refined below as the first block of the run
                    }
                    // ZRL: sixteen zero-history coefficients are skipped (value stays 0)
                } else {
                    if (size != 1) fail("bad refinement code");
                    value = bit() ? delta : -delta;
                }
                while (i <= se) {
                    int16_t& v = c[kNatural[i++]];
                    if (v != 0) {
                        refine(v);
                    } else {
                        if (run == 0) { v = int16_t(value); break; }
                        --run;
                    }
                }
            }
        }
        if (eob_run) {
            --eob_run;
            for (; i <= se; ++i) {
                int16_t& v = c[kNatural[i]];
                if (v != 0) refine(v);
            }
        }
    }

    // ---- inverse DCT (stb_image's integer arithmetic; see the header of this file) -------------------------------------------
    static constexpr int fix12(float x) { return int(x * 4096 + 0.5); }
    struct Butterfly { int e0, e1, e2, e3, o0, o1, o2, o3; };   // even part x0..x3, odd part t0..t3
    static Butterfly lines(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7) {
        Butterfly r;
        const int z = (s2 + s6) * fix12(0.5411961f);
        const int even_a = z + s6 * fix12(-1.847759065f), even_b = z + s2 * fix12(0.765366865f);
        const int sum = (s0 + s4) * 4096, diff = (s0 - s4) * 4096;
        r.e0 = sum + even_b; r.e3 = sum - even_b; r.e1 = diff + even_a; r.e2 = diff - even_a;
        const int a = s7 + s3, b = s5 + s1, c = s7 + s1, d = s5 + s3;
        const int z5 = (a + b) * fix12(1.175875602f);
        const int pc = z5 + c * fix12(-0.899976223f), pd = z5 + d * fix12(-2.562915447f);
        const int pa = a * fix12(-1.961570560f), pb = b * fix12(-0.390180644f);
        r.o3 = s1 * fix12(1.501321110f) + pc + pb;
        r.o2 = s3 * fix12(3.072711026f) + pd + pa;
        r.o1 = s5 * fix12(2.053119869f) + pd + pb;
        r.o0 = s7 * fix12(0.298631336f) + pc + pa;
        return r;
    }
    static uint8_t clamp8(int v) { return uint8_t(unsigned(v) > 255u ? (v < 0 ? 0 : 255) : v); }
    static void idct(const int16_t* d, uint8_t* out, int stride) {
        int mid[64];
        for (int i = 0; i < 8; ++i) {   // columns
            const Butterfly b = lines(d[i], d[8 + i], d[16 + i], d[24 + i], d[32 + i], d[40 + i], d[48 + i], d[56 + i]);
            const int e0 = b.e0 + 512, e1 = b.e1 + 512, e2 = b.e2 + 512, e3 = b.e3 + 512;
            mid[i] = (e0 + b.o3) >> 10; mid[56 + i] = (e0 - b.o3) >> 10;
            mid[8 + i] = (e1 + b.o2) >> 10; mid[48 + i] = (e1 - b.o2) >> 10;
            mid[16 + i] = (e2 + b.o1) >> 10; mid[40 + i] = (e2 - b.o1) >> 10;
            mid[24 + i] = (e3 + b.o0) >> 10; mid[32 + i] = (e3 - b.o0) >> 10;
        }
        for (int i = 0; i < 8; ++i) {   // rows; + 128 (level shift) and the rounding constant in one addend
            const int* v = mid + 8 * i;
            uint8_t* o = out + size_t(i) * stride;
            const Butterfly b = lines(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
            const int bias = 65536 + (128 << 17);
            const int e0 = b.e0 + bias, e1 = b.e1 + bias, e2 = b.e2 + bias, e3 = b.e3 + bias;
            o[0] = clamp8((e0 + b.o3) >> 17); o[7] = clamp8((e0 - b.o3) >> 17);
            o[1] = clamp8((e1 + b.o2) >> 17); o[6] = clamp8((e1 - b.o2) >> 17);
            o[2] = clamp8((e2 + b.o1) >> 17); o[5] = clamp8((e2 - b.o1) >> 17);
            o[3] = clamp8((e3 + b.o0) >> 17); o[4] = clamp8((e3 - b.o0) >> 17);
        }
    }

    // ---- a scan (T.81 A.2: interleaved scans walk MCUs, a one-component scan walks that component's own blocks) ----------------
    void decode_scan() {
        reset_entropy();
        int todo = restart_interval ? restart_interval : 0x7fffffff;
        int expected_rst = 0;
        auto after_unit = [&]() -> bool {   // false: the entropy-coded data ended early (a marker other than RSTn)
            if (--todo > 0) return true;
            if (bits < 24) fill();
            if (pending_marker < 0xd0 || pending_marker > 0xd7) return false;
            (void)expected_rst;
            reset_entropy();
            todo = restart_interval ? restart_interval : 0x7fffffff;
            return true;
        };
        int16_t block[64];
        if (scan_n == 1) {
            Component& k = comp[scan_comp[0]];
            const int w = (k.x + 7) >> 3, h = (k.y + 7) >> 3;
            for (int by = 0; by < h; ++by)
                for (int bx = 0; bx < w; ++bx) {
                    if (progressive) {
                        int16_t* c = k.coef.data() + (size_t(by) * k.blocks_w + bx) * 64;
                        if (ss == 0) block_progressive_dc(k, c); else block_progressive_ac(k, c);
                    } else {
                        block_sequential(k, block);
                        idct(block, k.plane.data() + (size_t(by) * 8 * k.blocks_w + bx) * 8, k.blocks_w * 8);
                    }
                    if (!after_unit()) return;
                }
        } else {
            for (int my = 0; my < mcus_y; ++my)
                for (int mx = 0; mx < mcus_x; ++mx) {
                    for (int i = 0; i < scan_n; ++i) {
                        Component& k = comp[scan_comp[i]];
                        for (int v = 0; v < k.v; ++v)
                            for (int u = 0; u < k.h; ++u) {
                                const int bx = mx * k.h + u, by = my * k.v + v;
                                if (progressive) {
                                    block_progressive_dc(k, k.coef.data() + (size_t(by) * k.blocks_w + bx) * 64);
                                } else {
                                    block_sequential(k, block);
                                    idct(block, k.plane.data() + (size_t(by) * 8 * k.blocks_w + bx) * 8, k.blocks_w * 8);
                                }
                            }
                    }
                    if (!after_unit()) return;
                }
        }
    }

    // progressive: all scans are in; dequantise and transform every block that lies inside the component
    void finish_progressive() {
        int16_t block[64];
        for (int c = 0; c < ncomp; ++c) {
            Component& k = comp[c];
            if (!quant_defined[k.tq]) fail("frame uses an undefined quantisation table");
            const uint16_t* q = quant[k.tq];
            const int w = (k.x + 7) >> 3, h = (k.y + 7) >> 3;
            for (int by = 0; by < h; ++by)
                for (int bx = 0; bx < w; ++bx) {
                    const int16_t* src = k.coef.data() + (size_t(by) * k.blocks_w + bx) * 64;
                    for (int i = 0; i < 64; ++i) block[i] = int16_t(src[i] * q[i]);
                    idct(block, k.plane.data() + (size_t(by) * 8 * k.blocks_w + bx) * 8, k.blocks_w * 8);
                }
        }
    }

    void parse() {
        if (n < 4 || p[0] != 0xff || p[1] != 0xd8) fail("not a JPEG file (no SOI marker)");
        pos = 2;
        bool seen_scan = false;
        for (;;) {
            int marker;
            if (pending_marker >= 0) {
                marker = pending_marker; pending_marker = -1;
            } else {
                // markers may be preceded by fill bytes; stray bytes after a scan's data are skipped up to the next marker
                uint8_t b = byte();
                while (b != 0xff) { if (pos >= n) { if (seen_scan) return; fail("file is truncated"); } b = byte(); }
                do { b = byte(); } while (b == 0xff);
                if (b == 0) continue;
                marker = b;
            }
            if (marker == 0xd9) break;                                   // EOI
            if (marker >= 0xd0 && marker <= 0xd7) continue;              // stray RSTn
            if (marker == 0x01) continue;                                // TEM
            const int len = be16();
            if (len < 2 || pos + size_t(len - 2) > n) fail("bad segment length");
            const size_t next = pos + size_t(len - 2);
            switch (marker) {
                case 0xdb: read_dqt(len - 2); break;
                case 0xc4: read_dht(len - 2); break;
                case 0xdd: if (len != 4) fail("bad DRI segment"); restart_interval = be16(); break;
                case 0xda:
                    read_sos(len);
                    pos = next;
                    decode_scan();
                    seen_scan = true;
                    continue;
                case 0xe0: if (len >= 7 && std::memcmp(p + pos, "JFIF", 5) == 0) jfif = true; break;
                case 0xee: if (len >= 14 && std::memcmp(p + pos, "Adobe", 6) == 0) adobe_transform = p[pos + 11]; break;
                case 0xdc: fail("DNL is not supported");
                default:
                    if (marker >= 0xc0 && marker <= 0xcf && marker != 0xc8) { read_sof(marker, len); break; }   // 0xc4 handled above
                    if ((marker >= 0xe0 && marker <= 0xef) || marker == 0xfe) break;                                  // APPn, COM
                    fail("unknown marker");
            }
            pos = next;
        }
        if (!ncomp || !seen_scan) fail("no image data");
        if (progressive) finish_progressive();
    }

    // ---- upsampling and colour conversion (stb_image's arithmetic; see the header of this file) ----------------------------------
    static void upsample_h2(uint8_t* out, const uint8_t* in, int w) {
        if (w == 1) { out[0] = out[1] = in[0]; return; }
        out[0] = in[0];
        out[1] = uint8_t((in[0] * 3 + in[1] + 2) >> 2);
        for (int i = 1; i < w - 1; ++i) {
            const int m = 3 * in[i] + 2;
            out[2 * i] = uint8_t((m + in[i - 1]) >> 2);
            out[2 * i + 1] = uint8_t((m + in[i + 1]) >> 2);
        }
        out[2 * w - 2] = uint8_t((in[w - 2] * 3 + in[w - 1] + 2) >> 2);   // stb_image weights the FAR sample here (not the mirror image of out[1]); kept, the bytes must match
        out[2 * w - 1] = in[w - 1];
    }
    static void upsample_v2(uint8_t* out, const uint8_t* near, const uint8_t* far, int w) {
        for (int i = 0; i < w; ++i) out[i] = uint8_t((3 * near[i] + far[i] + 2) >> 2);
    }
    static void upsample_h2v2(uint8_t* out, const uint8_t* near, const uint8_t* far, int w) {
        int cur = 3 * near[0] + far[0];
        if (w == 1) { out[0] = out[1] = uint8_t((cur + 2) >> 2); return; }
        out[0] = uint8_t((cur + 2) >> 2);
        for (int i = 1; i < w; ++i) {
            const int prev = cur;
            cur = 3 * near[i] + far[i];
            out[2 * i - 1] = uint8_t((3 * prev + cur + 8) >> 4);
            out[2 * i] = uint8_t((3 * cur + prev + 8) >> 4);
        }
        out[2 * w - 1] = uint8_t((cur + 2) >> 2);
    }

    JpegImage image() {
        JpegImage im;
        im.width = width; im.height = height; im.channels = ncomp;
        im.samples.resize(size_t(width) * height * ncomp);
        // is the 3-component image RGB already?  (stb_image: the Adobe transform flag, else the component identifiers 'R' 'G' 'B')
        bool rgb = false;
        if (ncomp == 3) rgb = (comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B') || (adobe_transform == 0 && !jfif);
        struct Up { int hs, vs, step, row, w_lores; const uint8_t *line0, *line1; std::vector<uint8_t> buf; } up[3];
        for (int c = 0; c < ncomp; ++c) {
            Up& u = up[c];
            u.hs = h_max / comp[c].h; u.vs = v_max / comp[c].v;
            u.step = u.vs >> 1; u.row = 0;
            u.w_lores = (width + u.hs - 1) / u.hs;
            u.line0 = u.line1 = comp[c].plane.data();
            u.buf.resize(size_t(width) + 3);
        }
        const uint8_t* rows[3] = {};
        for (int y = 0; y < height; ++y) {
            for (int c = 0; c < ncomp; ++c) {
                Up& u = up[c];
                const bool bottom = u.step >= (u.vs >> 1);   // which of the two source rows is nearer to this output row
                const uint8_t* near = bottom ? u.line1 : u.line0;
                const uint8_t* far = bottom ? u.line0 : u.line1;
                if (u.hs == 1 && u.vs == 1) rows[c] = near;
                else {
                    if (u.hs == 1) upsample_v2(u.buf.data(), near, far, u.w_lores);
                    else if (u.vs == 1) upsample_h2(u.buf.data(), near, u.w_lores);
                    else upsample_h2v2(u.buf.data(), near, far, u.w_lores);
                    rows[c] = u.buf.data();
                }
                if (++u.step >= u.vs) {
                    u.step = 0;
                    u.line0 = u.line1;
                    if (++u.row < comp[c].y) u.line1 += size_t(comp[c].blocks_w) * 8;
                }
            }
            uint8_t* dst = im.samples.data() + size_t(y) * width * ncomp;
            if (ncomp == 1) {
                std::memcpy(dst, rows[0], size_t(width));
            } else if (rgb) {
                for (int x = 0; x < width; ++x) { dst[3 * x] = rows[0][x]; dst[3 * x + 1] = rows[1][x]; dst[3 * x + 2] = rows[2][x]; }
            } else {
                constexpr auto fx = [](float v) { return int(v * 4096.0f + 0.5f) << 8; };
                for (int x = 0; x < width; ++x) {
                    const int luma = (rows[0][x] << 20) + (1 << 19);
                    const int cb = rows[1][x] - 128, cr = rows[2][x] - 128;
                    const int r = (luma + cr * fx(1.40200f)) >> 20;
                    const int g = (luma + cr * -fx(0.71414f) + int((unsigned(cb * -fx(0.34414f))) & 0xffff0000u)) >> 20;
                    const int b = (luma + cb * fx(1.77200f)) >> 20;
                    dst[3 * x] = clamp8(r); dst[3 * x + 1] = clamp8(g); dst[3 * x + 2] = clamp8(b);
                }
            }
        }
        return im;
    }
};

}  // namespace

JpegImage decode_jpeg(const uint8_t* data, size_t n) {
    Decoder d{data, n};
    d.parse();
    return d.image();
}

JpegImage load_jpeg(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("jpeg: cannot open " + path);
    std::vector<uint8_t> bytes;
    uint8_t chunk[65536];
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof(chunk), f)) > 0) bytes.insert(bytes.end(), chunk, chunk + got);
    std::fclose(f);
    return decode_jpeg(bytes.data(), bytes.size());
}

}  // namespace skyhost
