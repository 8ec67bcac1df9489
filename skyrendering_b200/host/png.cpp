// PNG reader for the reference's image inputs (the 64x64 16-bit blue-noise tile data/BlueNoise/64_64/HDR_L_0.png that Textures::Textures
// loads with stbi_load_16, src/Base/src/Textures.cpp:19-26).  stb_image is a vendored third-party header of the reference
// (external/stb); this restates the published formats it implements for this path -- RFC 2083 (PNG) and RFC 1950 / 1951 (zlib / DEFLATE) --
// from the specifications: non-interlaced images of colour type 0 / 2 / 4 / 6 (grey, RGB, grey + alpha, RGBA) at 8 or 16 bits per
// sample, all five scan-line filters, stored / fixed / dynamic Huffman blocks.  Everything else (palettes, sub-byte depths, Adam7) fails
// with a message.  Lossless, so parity is exact: tests compare against the shipped raw fixture and against files written by zlib.
#include "png.h"

#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace skyhost {
namespace {

struct BitReader {
    const uint8_t* p; size_t n, pos = 0; uint32_t acc = 0; int bits = 0;
    uint32_t take(int k) {   // k <= 16, LSB first (RFC 1951 3.1.1)
        while (bits < k) {
            if (pos >= n) throw std::runtime_error("png: deflate stream is truncated");
            acc |= uint32_t(p[pos++]) << bits; bits += 8;
        }
        uint32_t v = acc & ((1u << k) - 1u);
        acc >>= k; bits -= k;
        return v;
    }
    void align() { acc = 0; bits = 0; }
};

// canonical Huffman code (RFC 1951 3.2.2) decoded bit by bit through first-code / count tables
struct Huffman {
    uint16_t count[16] = {}, symbol[288] = {};
    void build(const uint8_t* lengths, int n) {
        std::memset(count, 0, sizeof(count));
        for (int i = 0; i < n; ++i) count[lengths[i]]++;
        count[0] = 0;
        uint16_t offs[16]; offs[1] = 0;
        for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + count[l];
        for (int i = 0; i < n; ++i) if (lengths[i]) symbol[offs[lengths[i]]++] = uint16_t(i);
    }
    int decode(BitReader& br) const {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len <= 15; ++len) {
            code |= int(br.take(1));
            int c = count[len];
            if (code - c < first) return symbol[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        throw std::runtime_error("png: bad Huffman code");
    }
};

std::vector<uint8_t> inflate(const uint8_t* data, size_t n, size_t expected) {
    if (n < 6) throw std::runtime_error("png: zlib stream is too short");
    if ((data[0] & 0x0f) != 8 || ((data[0] << 8) | data[1]) % 31 != 0 || (data[1] & 0x20)) throw std::runtime_error("png: bad zlib header");
    BitReader br{data + 2, n - 2};
    std::vector<uint8_t> out;
    out.reserve(expected);
    static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    for (bool last = false; !last;) {
        last = br.take(1) != 0;
        const uint32_t type = br.take(2);
        if (type == 0) {   // stored
            br.align();
            if (br.pos + 4 > br.n) throw std::runtime_error("png: deflate stream is truncated");
            const uint32_t len = br.p[br.pos] | (br.p[br.pos + 1] << 8), nlen = br.p[br.pos + 2] | (br.p[br.pos + 3] << 8);
            br.pos += 4;
            if ((len ^ 0xffffu) != nlen || br.pos + len > br.n) throw std::runtime_error("png: bad stored block");
            out.insert(out.end(), br.p + br.pos, br.p + br.pos + len);
            br.pos += len;
            continue;
        }
        if (type == 3) throw std::runtime_error("png: bad deflate block type");
        Huffman lit, dist;
        if (type == 1) {   // fixed codes, RFC 1951 3.2.6
            uint8_t l[288];
            for (int i = 0; i < 144; ++i) l[i] = 8;
            for (int i = 144; i < 256; ++i) l[i] = 9;
            for (int i = 256; i < 280; ++i) l[i] = 7;
            for (int i = 280; i < 288; ++i) l[i] = 8;
            lit.build(l, 288);
            uint8_t d[30];
            for (int i = 0; i < 30; ++i) d[i] = 5;
            dist.build(d, 30);
        } else {           // dynamic codes, 3.2.7
            const int hlit = int(br.take(5)) + 257, hdist = int(br.take(5)) + 1, hclen = int(br.take(4)) + 4;
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t cl[19] = {};
            for (int i = 0; i < hclen; ++i) cl[order[i]] = uint8_t(br.take(3));
            Huffman clh; clh.build(cl, 19);
            uint8_t lengths[288 + 32] = {};
            for (int i = 0; i < hlit + hdist;) {
                const int sym = clh.decode(br);
                if (sym < 16) { lengths[i++] = uint8_t(sym); continue; }
                int rep; uint8_t val = 0;
                if (sym == 16) { if (i == 0) throw std::runtime_error("png: bad code lengths"); val = lengths[i - 1]; rep = 3 + int(br.take(2)); }
                else if (sym == 17) rep = 3 + int(br.take(3));
                else rep = 11 + int(br.take(7));
                if (i + rep > hlit + hdist) throw std::runtime_error("png: bad code lengths");
                while (rep--) lengths[i++] = val;
            }
            if (hlit > 286 || hdist > 30) throw std::runtime_error("png: bad code counts");
            lit.build(lengths, hlit);
            dist.build(lengths + hlit, hdist);
        }
        for (;;) {
            const int sym = lit.decode(br);
            if (sym < 256) { out.push_back(uint8_t(sym)); continue; }
            if (sym == 256) break;
            if (sym > 285) throw std::runtime_error("png: bad length symbol");
            const int len = len_base[sym - 257] + int(br.take(len_extra[sym - 257]));
            const int ds = dist.decode(br);
            if (ds > 29) throw std::runtime_error("png: bad distance symbol");
            const size_t d = dist_base[ds] + br.take(dist_extra[ds]);
            if (d > out.size()) throw std::runtime_error("png: distance beyond the window");
            for (int k = 0; k < len; ++k) out.push_back(out[out.size() - d]);
        }
    }
    return out;
}

uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
int paeth(int a, int b, int c) {   // RFC 2083 6.6
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc) ? b : c;
}

}  // namespace

PngImage decode_png(const uint8_t* data, size_t n) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    if (n < 8 || std::memcmp(data, sig, 8) != 0) throw std::runtime_error("png: not a PNG file");
    PngImage im;
    std::vector<uint8_t> idat;
    bool have_header = false, done = false;
    int colour = 0, interlace = 0;
    for (size_t pos = 8; pos + 12 <= n && !done;) {
        const uint32_t len = be32(data + pos);
        if (pos + 12 + size_t(len) > n) throw std::runtime_error("png: chunk runs past the end of the file");
        const uint8_t* type = data + pos + 4; const uint8_t* body = data + pos + 8;
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) throw std::runtime_error("png: bad IHDR");
            im.width = int(be32(body)); im.height = int(be32(body + 4)); im.bits = body[8]; colour = body[9]; interlace = body[12];
            have_header = true;
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            done = true;
        }
        pos += 12 + size_t(len);
    }
    if (!have_header || idat.empty()) throw std::runtime_error("png: missing IHDR or IDAT");
    if (im.width <= 0 || im.height <= 0 || im.width > 32768 || im.height > 32768) throw std::runtime_error("png: bad dimensions");
    if (interlace != 0) throw std::runtime_error("png: Adam7 interlacing is not supported");
    if (im.bits != 8 && im.bits != 16) throw std::runtime_error("png: only 8 and 16 bits per sample are supported");
    switch (colour) { case 0: im.channels = 1; break; case 2: im.channels = 3; break; case 4: im.channels = 2; break; case 6: im.channels = 4; break;
                      default: throw std::runtime_error("png: palette images are not supported"); }
    const size_t bpp = size_t(im.channels) * (im.bits / 8), stride = size_t(im.width) * bpp;
    std::vector<uint8_t> raw = inflate(idat.data(), idat.size(), (stride + 1) * im.height);
    if (raw.size() < (stride + 1) * size_t(im.height)) throw std::runtime_error("png: image data is truncated");
    im.samples.resize(stride * im.height);   // big-endian samples as stored, rows top to bottom
    std::vector<uint8_t> zero(stride, 0);
    for (int y = 0; y < im.height; ++y) {
        const uint8_t* src = raw.data() + size_t(y) * (stride + 1);
        uint8_t* cur = im.samples.data() + size_t(y) * stride;
        const uint8_t* up = y ? cur - stride : zero.data();
        const int filter = src[0];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0, x = src[1 + i];
            int v;
            switch (filter) {
                case 0: v = x; break;
                case 1: v = x + a; break;
                case 2: v = x + b; break;
                case 3: v = x + ((a + b) >> 1); break;
                case 4: v = x + paeth(a, b, c); break;
                default: throw std::runtime_error("png: bad filter type");
            }
            cur[i] = uint8_t(v);
        }
    }
    return im;
}

PngImage load_png(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("png: cannot open " + path);
    std::vector<uint8_t> bytes;
    uint8_t buf[65536];
    for (size_t k; (k = std::fread(buf, 1, sizeof(buf), f)) > 0;) bytes.insert(bytes.end(), buf, buf + k);
    std::fclose(f);
    return decode_png(bytes.data(), bytes.size());
}

}  // namespace skyhost
