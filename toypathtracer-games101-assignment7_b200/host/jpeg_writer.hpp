// jpeg_writer.hpp — baseline sequential JPEG (ITU-T T.81) writer for 8-bit RGB frames.
//
// The reference writes its result with stbi_write_jpg(..., quality 100)
// (SceneRenderingHelper.cpp:68): YCbCr, no chroma subsampling, quantisation tables of all ones.
// This is an independent implementation of that file format (JFIF header, 8x8 forward DCT, the
// typical Huffman tables of T.81 Annex K.3) so that Renderer::Render("output.jpg", ...) produces a
// real .jpg like the reference does.  Host-side post-processing, not on the hot path.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace tptjpeg {

struct BitWriter {
    std::vector<unsigned char>& out;
    uint32_t acc = 0;
    int nbits = 0;
    explicit BitWriter(std::vector<unsigned char>& o) : out(o) {}
    void put(uint32_t code, int len) {
        acc = (acc << len) | (code & ((1u << len) - 1u));
        nbits += len;
        while (nbits >= 8) {
            const unsigned char byte = (unsigned char)((acc >> (nbits - 8)) & 0xffu);
            out.push_back(byte);
            if (byte == 0xff) out.push_back(0);      // byte stuffing, T.81 B.1.1.5
            nbits -= 8;
        }
    }
    void flush() { if (nbits > 0) put(0x7f, 8 - nbits); }     // pad with ones
};

struct Huffman {
    uint16_t code[256];
    uint8_t len[256];
    // bits[i] = number of codes of length i + 1; vals in order of increasing code (T.81 Annex C)
    Huffman(const uint8_t* bits, const uint8_t* vals) {
        for (int i = 0; i < 256; ++i) { code[i] = 0; len[i] = 0; }
        uint16_t c = 0;
        int k = 0;
        for (int l = 1; l <= 16; ++l) {
            for (int i = 0; i < bits[l - 1]; ++i, ++k) { code[vals[k]] = c++; len[vals[k]] = (uint8_t)l; }
            c <<= 1;
        }
    }
};

// T.81 Annex K.3: typical Huffman tables for 8-bit luminance / chrominance
static const uint8_t kDcLumBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t kDcChrBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kAcLumBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t kAcLumVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1,
    0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56,
    0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85,
    0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa,
    0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};
static const uint8_t kAcChrBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t kAcChrVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
    0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19,
    0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55,
    0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8,
    0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4,
    0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa};
// zig-zag order: kZigzag[k] = raster index of the k-th coefficient
static const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

inline void put16(std::vector<unsigned char>& o, int v) { o.push_back((unsigned char)(v >> 8)); o.push_back((unsigned char)(v & 0xff)); }
inline void put_dht(std::vector<unsigned char>& o, int tc_th, const uint8_t* bits, const uint8_t* vals, int nvals) {
    o.push_back(0xff); o.push_back(0xc4);
    put16(o, 2 + 1 + 16 + nvals);
    o.push_back((unsigned char)tc_th);
    o.insert(o.end(), bits, bits + 16);
    o.insert(o.end(), vals, vals + nvals);
}

// 8x8 forward DCT (separable, double precision; this is a file writer, not a hot loop)
inline void fdct8x8(const float* in, int* outq) {
    static double c[8][8];
    static bool init = false;
    if (!init) {
        for (int u = 0; u < 8; ++u)
            for (int x = 0; x < 8; ++x) c[u][x] = (u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * 3.14159265358979323846 / 16.0);
        init = true;
    }
    double tmp[64];
    for (int y = 0; y < 8; ++y)
        for (int u = 0; u < 8; ++u) {
            double s = 0;
            for (int x = 0; x < 8; ++x) s += c[u][x] * in[y * 8 + x];
            tmp[y * 8 + u] = s;
        }
    for (int u = 0; u < 8; ++u)
        for (int v = 0; v < 8; ++v) {
            double s = 0;
            for (int y = 0; y < 8; ++y) s += c[v][y] * tmp[y * 8 + u];
            outq[v * 8 + u] = (int)std::lround(s);           // quantisation table of ones (quality 100)
        }
}

inline void encode_block(BitWriter& bw, const int* q, int& pred, const Huffman& dc, const Huffman& ac) {
    auto category = [](int v, int& bits) {
        int a = v < 0 ? -v : v, n = 0;
        while (a) { ++n; a >>= 1; }
        bits = v < 0 ? v - 1 : v;          // low n bits of (v - 1) for negatives, T.81 F.1.2.1
        return n;
    };
    int bits;
    const int dcv = q[0] < -1024 ? -1024 : (q[0] > 1023 ? 1023 : q[0]);
    const int diff = dcv - pred;
    pred = dcv;
    int n = category(diff, bits);
    bw.put(dc.code[n], dc.len[n]);
    if (n) bw.put((uint32_t)bits, n);
    int run = 0;
    for (int k = 1; k < 64; ++k) {
        int v = q[kZigzag[k]];
        v = v < -1023 ? -1023 : (v > 1023 ? 1023 : v);
        if (v == 0) { ++run; continue; }
        while (run > 15) { bw.put(ac.code[0xf0], ac.len[0xf0]); run -= 16; }
        n = category(v, bits);
        const int sym = (run << 4) | n;
        bw.put(ac.code[sym], ac.len[sym]);
        bw.put((uint32_t)bits, n);
        run = 0;
    }
    if (run) bw.put(ac.code[0x00], ac.len[0x00]);          // EOB
}

// rgb: width*height*3 bytes, row-major, top row first.  Returns false if the file cannot be written.
inline bool write_jpeg(const std::string& path, const unsigned char* rgb, int width, int height) {
    std::vector<unsigned char> o;
    o.reserve((size_t)width * height);
    o.push_back(0xff); o.push_back(0xd8);                                   // SOI
    const unsigned char jfif[] = {0xff, 0xe0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
    o.insert(o.end(), jfif, jfif + sizeof jfif);
    for (int t = 0; t < 2; ++t) {                                           // DQT: all ones
        o.push_back(0xff); o.push_back(0xdb); put16(o, 67); o.push_back((unsigned char)t);
        for (int i = 0; i < 64; ++i) o.push_back(1);
    }
    o.push_back(0xff); o.push_back(0xc0); put16(o, 17); o.push_back(8);     // SOF0, 8 bit
    put16(o, height); put16(o, width); o.push_back(3);
    const unsigned char comps[9] = {1, 0x11, 0, 2, 0x11, 1, 3, 0x11, 1};    // Y, Cb, Cr: 1x1 sampling (4:4:4)
    o.insert(o.end(), comps, comps + 9);
    put_dht(o, 0x00, kDcLumBits, kDcVals, 12);
    put_dht(o, 0x10, kAcLumBits, kAcLumVals, 162);
    put_dht(o, 0x01, kDcChrBits, kDcVals, 12);
    put_dht(o, 0x11, kAcChrBits, kAcChrVals, 162);
    const unsigned char sos[] = {0xff, 0xda, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
    o.insert(o.end(), sos, sos + sizeof sos);

    const Huffman dcL(kDcLumBits, kDcVals), acL(kAcLumBits, kAcLumVals), dcC(kDcChrBits, kDcVals), acC(kAcChrBits, kAcChrVals);
    BitWriter bw(o);
    int pred[3] = {0, 0, 0};
    float blk[3][64];
    int q[64];
    for (int by = 0; by < height; by += 8)
        for (int bx = 0; bx < width; bx += 8) {
            for (int y = 0; y < 8; ++y)
                for (int x = 0; x < 8; ++x) {
                    const int px = bx + x < width ? bx + x : width - 1, py = by + y < height ? by + y : height - 1;   // edge replication
                    const unsigned char* p = rgb + 3 * ((size_t)py * width + px);
                    const float r = p[0], g = p[1], b = p[2];
                    blk[0][y * 8 + x] = 0.299f * r + 0.587f * g + 0.114f * b - 128.0f;
                    blk[1][y * 8 + x] = -0.168736f * r - 0.331264f * g + 0.5f * b;
                    blk[2][y * 8 + x] = 0.5f * r - 0.418688f * g - 0.081312f * b;
                }
            for (int c = 0; c < 3; ++c) {
                fdct8x8(blk[c], q);
                encode_block(bw, q, pred[c], c == 0 ? dcL : dcC, c == 0 ? acL : acC);
            }
        }
    bw.flush();
    o.push_back(0xff); o.push_back(0xd9);                                   // EOI
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(o.data(), 1, o.size(), f) == o.size();
    std::fclose(f);
    return ok;
}

}  // namespace tptjpeg
