// See image_decode.hpp.
#include "image_decode.hpp"

#include <algorithm>
#include <climits>
#include <cstring>

namespace cray {
namespace {

// ================================================================ JPEG (ITU T.81) ===============================================

// zigzag position -> natural (row-major) position, T.81 figure A.6
const uint8_t kNatural[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                              41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                              30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct JpegError {
    std::string message;
};

// Canonical Huffman code of one DHT table (T.81 Annex C), decoded with a 9-bit first-level table and the
// mincode / maxcode / valptr search of F.2.2.3 for the longer codes.
struct HuffTable {
    bool present = false;
    uint8_t counts[17] = {};
    uint8_t symbols[256] = {};
    int32_t mincode[17] = {}, maxcode[18] = {}, valptr[17] = {};
    uint16_t fast[512];  // length << 8 | symbol, 0xFFFF: longer than 9 bits

    void build() {
        int32_t code = 0, k = 0;
        for (int len = 1; len <= 16; ++len) {
            valptr[len] = k;
            mincode[len] = code;
            code += counts[len];
            k += counts[len];
            maxcode[len] = counts[len] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = INT_MAX;
        std::fill(fast, fast + 512, (uint16_t)0xFFFF);
        code = 0;
        k = 0;
        for (int len = 1; len <= 9; ++len) {
            for (int i = 0; i < counts[len]; ++i, ++k, ++code) {
                const int first = code << (9 - len), n = 1 << (9 - len);
                for (int j = 0; j < n && first + j < 512; ++j) fast[first + j] = (uint16_t)(len << 8 | symbols[k]);
            }
            code <<= 1;
        }
        present = true;
    }
};

// Entropy-coded segment reader: removes the stuffed zero after 0xFF and stops at a marker (zeros are supplied from there on).
struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint32_t acc = 0;
    int count = 0;
    bool at_marker = false;

    void fill() {
        while (count <= 24) {
            uint32_t byte = 0;
            if (!at_marker && p < end) {
                byte = *p;
                if (byte == 0xFF) {
                    const uint8_t next = p + 1 < end ? p[1] : 0xD9;
                    if (next == 0x00) p += 2;
                    else if (next == 0xFF) { ++p; continue; }  // fill byte
                    else { at_marker = true; byte = 0; }
                } else {
                    ++p;
                }
            }
            acc |= byte << (24 - count);
            count += 8;
        }
    }
    uint32_t peek(int n) {
        if (count < n) fill();
        return acc >> (32 - n);
    }
    void skip(int n) { acc <<= n; count -= n; }
    int bits(int n) {
        if (n == 0) return 0;
        const uint32_t v = peek(n);
        skip(n);
        return (int)v;
    }
    int bit() { return bits(1); }
    // discards the rest of the current byte and steps over the RSTn marker that must follow
    void restart() {
        acc = 0;
        count = 0;
        at_marker = false;
        while (p + 1 < end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) {
            if (p[0] == 0xFF && p[1] != 0x00 && p[1] != 0xFF) return;  // some other marker: leave it to the segment parser
            ++p;
        }
        if (p + 1 < end) p += 2;
    }
    int decode(const HuffTable& t) {
        const uint32_t look = peek(16);
        const uint16_t f = t.fast[look >> 7];
        if (f != 0xFFFF) { skip(f >> 8); return f & 0xFF; }
        int len = 10;
        int32_t code = (int32_t)(look >> 6);
        while (len <= 16 && code > t.maxcode[len]) { ++len; code = (int32_t)(look >> (16 - len)); }
        if (len > 16) throw JpegError{"corrupt JPEG: bad Huffman code"};
        skip(len);
        return t.symbols[t.valptr[len] + code - t.mincode[len]];
    }
};

inline int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }  // T.81 F.2.2.1

struct Component {
    int id = 0, h = 1, v = 1, tq = 0;
    int blocks_w = 0, blocks_h = 0;   // allocated: whole MCUs
    int scan_w = 0, scan_h = 0;       // blocks of a non-interleaved scan: ceil(ceil(W * h / hmax) / 8)
    int ds_w = 0, ds_h = 0;           // samples that belong to the image (libjpeg's downsampled_width / height)
    int dc_pred = 0, td = 0, ta = 0;
    std::vector<int16_t> coef;        // blocks_w * blocks_h * 64, natural order
    std::vector<uint8_t> plane;       // blocks_w * 8 x blocks_h * 8
};

struct Decoder {
    const uint8_t* data;
    size_t n;
    Decoder(const uint8_t* bytes, size_t size) : data(bytes), n(size) {}
    size_t pos = 0;
    uint16_t quant[4][64] = {};
    bool quant_present[4] = {};
    HuffTable dc_tables[4], ac_tables[4];
    std::vector<Component> comps;
    int width = 0, height = 0, hmax = 1, vmax = 1, mcus_x = 0, mcus_y = 0;
    bool progressive = false, have_frame = false;
    int restart_interval = 0;
    bool jfif = false, adobe = false;
    int adobe_transform = 0;

    uint8_t u8() {
        if (pos >= n) throw JpegError{"truncated JPEG"};
        return data[pos++];
    }
    int u16() { const int a = u8(); return a << 8 | u8(); }

    void read_dqt(size_t end) {
        while (pos < end) {
            const int pq_tq = u8(), pq = pq_tq >> 4, tq = pq_tq & 15;
            if (tq > 3 || pq > 1) throw JpegError{"corrupt JPEG: bad quantisation table"};
            for (int i = 0; i < 64; ++i) quant[tq][kNatural[i]] = (uint16_t)(pq ? u16() : u8());
            quant_present[tq] = true;
        }
    }
    void read_dht(size_t end) {
        while (pos < end) {
            const int tc_th = u8(), tc = tc_th >> 4, th = tc_th & 15;
            if (tc > 1 || th > 3) throw JpegError{"corrupt JPEG: bad Huffman table id"};
            HuffTable& t = tc ? ac_tables[th] : dc_tables[th];
            int total = 0;
            t.counts[0] = 0;
            for (int i = 1; i <= 16; ++i) { t.counts[i] = u8(); total += t.counts[i]; }
            if (total > 256) throw JpegError{"corrupt JPEG: bad Huffman table"};
            for (int i = 0; i < total; ++i) t.symbols[i] = u8();
            t.build();
        }
    }
    void read_sof(int marker) {
        if (have_frame) throw JpegError{"unsupported JPEG: more than one frame"};
        progressive = marker == 0xC2;
        const int precision = u8();
        height = u16();
        width = u16();
        const int nc = u8();
        if (precision != 8) throw JpegError{"unsupported JPEG: sample precision is not 8 bits"};
        if (width <= 0 || height <= 0) throw JpegError{"unsupported JPEG: empty frame"};
        if (nc != 1 && nc != 3) throw JpegError{"unsupported JPEG: " + std::to_string(nc) + " components"};
        comps.resize(nc);
        for (Component& c : comps) {
            c.id = u8();
            const int hv = u8();
            c.h = hv >> 4; c.v = hv & 15; c.tq = u8();
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) throw JpegError{"corrupt JPEG: bad component"};
            hmax = std::max(hmax, c.h); vmax = std::max(vmax, c.v);
        }
        if (nc == 1) { comps[0].h = comps[0].v = 1; hmax = vmax = 1; }  // a single component is never subsampled (A.2.2)
        mcus_x = (width + 8 * hmax - 1) / (8 * hmax);
        mcus_y = (height + 8 * vmax - 1) / (8 * vmax);
        for (Component& c : comps) {
            c.blocks_w = mcus_x * c.h; c.blocks_h = mcus_y * c.v;
            c.ds_w = (width * c.h + hmax - 1) / hmax; c.ds_h = (height * c.v + vmax - 1) / vmax;
            c.scan_w = (c.ds_w + 7) / 8; c.scan_h = (c.ds_h + 7) / 8;
            c.coef.assign((size_t)c.blocks_w * c.blocks_h * 64, 0);
        }
        have_frame = true;
    }

    // ---- one block of each scan kind (T.81 F.2.2, G.1.2) ----
    static void block_baseline(BitReader& br, Component& c, const HuffTable& dc, const HuffTable& ac, int16_t* blk) {
        const int t = br.decode(dc);
        if (t > 11) throw JpegError{"corrupt JPEG: bad DC size"};
        c.dc_pred += t ? extend(br.bits(t), t) : 0;
        blk[0] = (int16_t)c.dc_pred;
        for (int k = 1; k < 64;) {
            const int rs = br.decode(ac), r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r != 15) break;
                k += 16;
                continue;
            }
            k += r;
            if (k > 63) throw JpegError{"corrupt JPEG: coefficient index out of range"};
            blk[kNatural[k]] = (int16_t)extend(br.bits(s), s);
            ++k;
        }
    }
    static void block_dc_first(BitReader& br, Component& c, const HuffTable& dc, int16_t* blk, int al) {
        const int t = br.decode(dc);
        if (t > 11) throw JpegError{"corrupt JPEG: bad DC size"};
        c.dc_pred += t ? extend(br.bits(t), t) : 0;
        blk[0] = (int16_t)(c.dc_pred * (1 << al));
    }
    static void block_dc_refine(BitReader& br, int16_t* blk, int al) {
        if (br.bit()) blk[0] |= (int16_t)(1 << al);
    }
    static void block_ac_first(BitReader& br, const HuffTable& ac, int16_t* blk, int ss, int se, int al, int& eobrun) {
        if (eobrun > 0) { --eobrun; return; }
        for (int k = ss; k <= se;) {
            const int rs = br.decode(ac), r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (r < 15) {
                    eobrun = (1 << r) - 1;
                    if (r) eobrun += br.bits(r);
                    break;
                }
                k += 16;
            } else {
                k += r;
                if (k > 63) throw JpegError{"corrupt JPEG: coefficient index out of range"};
                blk[kNatural[k]] = (int16_t)(extend(br.bits(s), s) * (1 << al));
                ++k;
            }
        }
    }
    static void block_ac_refine(BitReader& br, const HuffTable& ac, int16_t* blk, int ss, int se, int al, int& eobrun) {
        const int p1 = 1 << al, m1 = -(1 << al);
        int k = ss;
        auto refine = [&](int16_t& c) {
            if (br.bit() && (c & p1) == 0) c = (int16_t)(c + (c >= 0 ? p1 : m1));
        };
        if (eobrun <= 0) {
            for (; k <= se; ++k) {
                const int rs = br.decode(ac);
                int r = rs >> 4, s = rs & 15;
                if (s) {
                    s = br.bit() ? p1 : m1;
                } else if (r != 15) {
                    eobrun = 1 << r;
                    if (r) eobrun += br.bits(r);
                    break;
                }
                // skip r zero-history coefficients, refining the nonzero ones on the way
                do {
                    int16_t& c = blk[kNatural[k]];
                    if (c != 0) refine(c);
                    else if (--r < 0) break;
                    ++k;
                } while (k <= se);
                if (s) {
                    if (k > 63) throw JpegError{"corrupt JPEG: coefficient index out of range"};
                    blk[kNatural[k]] = (int16_t)s;
                }
            }
        }
        if (eobrun > 0) {
            for (; k <= se; ++k) {
                int16_t& c = blk[kNatural[k]];
                if (c != 0) refine(c);
            }
            --eobrun;
        }
    }

    void read_scan() {
        if (!have_frame) throw JpegError{"corrupt JPEG: scan before frame header"};
        const int ns = u8();
        if (ns < 1 || ns > (int)comps.size()) throw JpegError{"corrupt JPEG: bad scan component count"};
        Component* sc[4] = {};
        for (int i = 0; i < ns; ++i) {
            const int cs = u8(), tdta = u8();
            for (Component& c : comps)
                if (c.id == cs) sc[i] = &c;
            if (!sc[i]) throw JpegError{"corrupt JPEG: scan names an unknown component"};
            sc[i]->td = tdta >> 4; sc[i]->ta = tdta & 15;
            if (sc[i]->td > 3 || sc[i]->ta > 3) throw JpegError{"corrupt JPEG: bad table selector"};
        }
        const int ss = u8(), se = u8(), ahal = u8(), ah = ahal >> 4, al = ahal & 15;
        if (progressive) {
            if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss > 0 && ns != 1) || al > 13) throw JpegError{"corrupt JPEG: bad progressive scan parameters"};
        } else if (ss != 0 || se != 63 || ah != 0 || al != 0) {
            throw JpegError{"corrupt JPEG: bad sequential scan parameters"};
        }
        for (int i = 0; i < ns; ++i) {
            const bool need_dc = !progressive || (ss == 0 && ah == 0), need_ac = !progressive || ss > 0;
            if (need_dc && !dc_tables[sc[i]->td].present) throw JpegError{"corrupt JPEG: missing DC Huffman table"};
            if (need_ac && !ac_tables[sc[i]->ta].present) throw JpegError{"corrupt JPEG: missing AC Huffman table"};
            sc[i]->dc_pred = 0;
        }
        BitReader br{data + pos, data + n};
        int eobrun = 0;
        auto one_block = [&](Component& c, int bx, int by) {
            int16_t* blk = c.coef.data() + ((size_t)by * c.blocks_w + bx) * 64;
            if (!progressive) block_baseline(br, c, dc_tables[c.td], ac_tables[c.ta], blk);
            else if (ss == 0) { if (ah == 0) block_dc_first(br, c, dc_tables[c.td], blk, al); else block_dc_refine(br, blk, al); }
            else if (ah == 0) block_ac_first(br, ac_tables[c.ta], blk, ss, se, al, eobrun);
            else block_ac_refine(br, ac_tables[c.ta], blk, ss, se, al, eobrun);
        };
        const int units_x = ns == 1 ? sc[0]->scan_w : mcus_x, units_y = ns == 1 ? sc[0]->scan_h : mcus_y;
        int until_restart = restart_interval;
        for (int uy = 0; uy < units_y; ++uy)
            for (int ux = 0; ux < units_x; ++ux) {
                if (restart_interval && until_restart == 0) {
                    br.restart();
                    for (int i = 0; i < ns; ++i) sc[i]->dc_pred = 0;
                    eobrun = 0;
                    until_restart = restart_interval;
                }
                if (ns == 1) {
                    one_block(*sc[0], ux, uy);
                } else {
                    for (int i = 0; i < ns; ++i)
                        for (int vy = 0; vy < sc[i]->v; ++vy)
                            for (int hx = 0; hx < sc[i]->h; ++hx) one_block(*sc[i], ux * sc[i]->h + hx, uy * sc[i]->v + vy);
                }
                --until_restart;
            }
        // the next marker: where the reader stopped, or a little further if the encoder padded
        pos = (size_t)(br.p - data);
        if (!br.at_marker) {
            while (pos + 1 < n && !(data[pos] == 0xFF && data[pos + 1] != 0x00 && data[pos + 1] != 0xFF)) ++pos;
        }
    }

    // ---- inverse DCT: the IJG "slow integer" transform (Loeffler, Ligtenberg, Moschytz), 13-bit constants ----
    static inline int32_t descale(int32_t x, int n) { return (x + (1 << (n - 1))) >> n; }
    static inline uint8_t clamp8(int32_t v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

    static void idct_block(const int16_t* in, const uint16_t* q, uint8_t* out, size_t stride) {
        constexpr int32_t F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633, F1_501 = 12299,
                          F1_847 = 15137, F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172;
        constexpr int kConst = 13, kPass1 = 2;
        int32_t ws[64];
        auto butterfly = [&](int32_t d0, int32_t d1, int32_t d2, int32_t d3, int32_t d4, int32_t d5, int32_t d6, int32_t d7, int32_t* o) {
            // even part
            int32_t z1 = (d2 + d6) * F0_541;
            const int32_t t2 = z1 + d6 * (-F1_847), t3 = z1 + d2 * F0_765;
            const int32_t t0 = (d0 + d4) * (1 << kConst), t1 = (d0 - d4) * (1 << kConst);
            const int32_t t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
            // odd part
            int32_t o0 = d7, o1 = d5, o2 = d3, o3 = d1;
            z1 = o0 + o3;
            int32_t z2 = o1 + o2, z3 = o0 + o2, z4 = o1 + o3;
            const int32_t z5 = (z3 + z4) * F1_175;
            o0 *= F0_298; o1 *= F2_053; o2 *= F3_072; o3 *= F1_501;
            z1 *= -F0_899; z2 *= -F2_562; z3 *= -F1_961; z4 *= -F0_390;
            z3 += z5; z4 += z5;
            o0 += z1 + z3; o1 += z2 + z4; o2 += z2 + z3; o3 += z1 + z4;
            o[0] = t10 + o3; o[7] = t10 - o3; o[1] = t11 + o2; o[6] = t11 - o2;
            o[2] = t12 + o1; o[5] = t12 - o1; o[3] = t13 + o0; o[4] = t13 - o0;
        };
        for (int c = 0; c < 8; ++c) {  // columns
            int32_t o[8];
            butterfly(in[c] * q[c], in[8 + c] * q[8 + c], in[16 + c] * q[16 + c], in[24 + c] * q[24 + c], in[32 + c] * q[32 + c], in[40 + c] * q[40 + c],
                      in[48 + c] * q[48 + c], in[56 + c] * q[56 + c], o);
            for (int r = 0; r < 8; ++r) ws[r * 8 + c] = descale(o[r], kConst - kPass1);
        }
        for (int r = 0; r < 8; ++r) {  // rows
            const int32_t* w = ws + r * 8;
            int32_t o[8];
            butterfly(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], o);
            for (int c = 0; c < 8; ++c) out[r * stride + c] = clamp8(descale(o[c], kConst + kPass1 + 3) + 128);
        }
    }

    void reconstruct_planes() {
        for (Component& c : comps) {
            if (!quant_present[c.tq]) throw JpegError{"corrupt JPEG: missing quantisation table"};
            const size_t stride = (size_t)c.blocks_w * 8;
            c.plane.assign(stride * c.blocks_h * 8, 0);
            for (int by = 0; by < c.blocks_h; ++by)
                for (int bx = 0; bx < c.blocks_w; ++bx)
                    idct_block(c.coef.data() + ((size_t)by * c.blocks_w + bx) * 64, quant[c.tq], c.plane.data() + (size_t)by * 8 * stride + (size_t)bx * 8, stride);
            c.coef.clear();
            c.coef.shrink_to_fit();
        }
    }

    // One full-resolution row `y` of component c (IJG upsampling: triangle filters for 2:1 horizontally and / or vertically when the
    // component is more than two samples wide, sample replication otherwise).
    void upsampled_row(const Component& c, int y, uint8_t* out) const {
        const size_t stride = (size_t)c.blocks_w * 8;
        const int fx = hmax / c.h, fy = vmax / c.v;
        const bool exact = hmax % c.h == 0 && vmax % c.v == 0;
        const int n = c.ds_w;
        auto row = [&](int r) { return c.plane.data() + (size_t)std::min(std::max(r, 0), c.ds_h - 1) * stride; };
        if (fx == 1 && fy == 1) {
            std::memcpy(out, row(y), (size_t)width);
            return;
        }
        std::vector<uint8_t> wide;  // 2 * n samples of a horizontally doubled row
        if (exact && fx == 2 && fy == 2 && n > 2) {
            const int r = y / 2;
            const uint8_t *near = row(r), *far = row((y & 1) ? r + 1 : r - 1);
            wide.resize((size_t)2 * n);
            int this_sum = near[0] * 3 + far[0], next_sum = near[1] * 3 + far[1], last_sum;
            wide[0] = (uint8_t)((this_sum * 4 + 8) >> 4);
            wide[1] = (uint8_t)((this_sum * 3 + next_sum + 7) >> 4);
            for (int i = 1; i < n - 1; ++i) {
                last_sum = this_sum; this_sum = next_sum;
                next_sum = near[i + 1] * 3 + far[i + 1];
                wide[2 * i] = (uint8_t)((this_sum * 3 + last_sum + 8) >> 4);
                wide[2 * i + 1] = (uint8_t)((this_sum * 3 + next_sum + 7) >> 4);
            }
            last_sum = this_sum; this_sum = next_sum;
            wide[2 * n - 2] = (uint8_t)((this_sum * 3 + last_sum + 8) >> 4);
            wide[2 * n - 1] = (uint8_t)((this_sum * 4 + 7) >> 4);
            std::memcpy(out, wide.data(), (size_t)width);
            return;
        }
        if (exact && fx == 2 && fy == 1 && n > 2) {
            const uint8_t* in = row(y);
            wide.resize((size_t)2 * n);
            wide[0] = in[0];
            wide[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
            for (int i = 1; i < n - 1; ++i) {
                wide[2 * i] = (uint8_t)((in[i] * 3 + in[i - 1] + 1) >> 2);
                wide[2 * i + 1] = (uint8_t)((in[i] * 3 + in[i + 1] + 2) >> 2);
            }
            wide[2 * n - 2] = (uint8_t)((in[n - 1] * 3 + in[n - 2] + 1) >> 2);
            wide[2 * n - 1] = in[n - 1];
            std::memcpy(out, wide.data(), (size_t)width);
            return;
        }
        if (exact && fx == 1 && fy == 2) {
            const int r = y / 2;
            const uint8_t *near = row(r), *far = row((y & 1) ? r + 1 : r - 1);
            const int bias = (y & 1) ? 2 : 1;
            for (int x = 0; x < width; ++x) out[x] = (uint8_t)((near[x] * 3 + far[x] + bias) >> 2);
            return;
        }
        // any other ratio: nearest sample
        const uint8_t* in = row(exact ? y / fy : (int)((int64_t)y * c.v / vmax));
        for (int x = 0; x < width; ++x) out[x] = in[std::min(exact ? x / fx : (int)((int64_t)x * c.h / hmax), n - 1)];
    }

    void to_rgb(std::vector<uint8_t>& rgb) const {
        rgb.resize((size_t)width * height * 3);
        if (comps.size() == 1) {
            const size_t stride = (size_t)comps[0].blocks_w * 8;
            for (int y = 0; y < height; ++y)
                for (int x = 0; x < width; ++x) {
                    const uint8_t g = comps[0].plane[(size_t)y * stride + x];
                    uint8_t* px = &rgb[((size_t)y * width + x) * 3];
                    px[0] = px[1] = px[2] = g;
                }
            return;
        }
        // colour space of a three-component file: JFIF says YCbCr, an Adobe marker says by its transform flag, otherwise by the ids
        bool ycc = true;
        if (jfif) ycc = true;
        else if (adobe) ycc = adobe_transform != 0;
        else if (comps[0].id == 'R' && comps[1].id == 'G' && comps[2].id == 'B') ycc = false;
        // fixed-point BT.601 tables, 16 fractional bits
        int32_t cr_r[256], cb_b[256], cr_g[256], cb_g[256];
        for (int i = 0; i < 256; ++i) {
            const int32_t x = i - 128;
            cr_r[i] = (91881 * x + 32768) >> 16;     // 1.40200
            cb_b[i] = (116130 * x + 32768) >> 16;    // 1.77200
            cr_g[i] = -46802 * x;                    // 0.71414
            cb_g[i] = -22554 * x + 32768;            // 0.34414
        }
        std::vector<uint8_t> r0((size_t)width), r1((size_t)width), r2((size_t)width);
        for (int y = 0; y < height; ++y) {
            upsampled_row(comps[0], y, r0.data());
            upsampled_row(comps[1], y, r1.data());
            upsampled_row(comps[2], y, r2.data());
            uint8_t* px = &rgb[(size_t)y * width * 3];
            for (int x = 0; x < width; ++x, px += 3) {
                if (ycc) {
                    const int Y = r0[x], cb = r1[x], cr = r2[x];
                    px[0] = clamp8(Y + cr_r[cr]);
                    px[1] = clamp8(Y + ((cb_g[cb] + cr_g[cr]) >> 16));
                    px[2] = clamp8(Y + cb_b[cb]);
                } else {
                    px[0] = r0[x]; px[1] = r1[x]; px[2] = r2[x];
                }
            }
        }
    }

    void run(uint32_t& w, uint32_t& h, std::vector<uint8_t>& rgb) {
        if (n < 4 || data[0] != 0xFF || data[1] != 0xD8) throw JpegError{"not a JPEG file"};
        pos = 2;
        bool done = false, any_scan = false;
        while (!done) {
            // next marker
            if (pos + 1 >= n) break;  // (a file cut off after its last scan still decodes, like the IJG decoder's premature-EOF path)
            if (data[pos] != 0xFF) { ++pos; continue; }
            const int marker = data[pos + 1];
            if (marker == 0xFF) { ++pos; continue; }
            pos += 2;
            if (marker == 0x00 || marker == 0x01 || (marker >= 0xD0 && marker <= 0xD7)) continue;
            if (marker == 0xD9) { done = true; break; }
            const size_t seg = pos;
            const int len = u16();
            if (len < 2 || seg + len > n) throw JpegError{"truncated JPEG"};
            const size_t end = seg + len;
            switch (marker) {
                case 0xDB: read_dqt(end); break;
                case 0xC4: read_dht(end); break;
                case 0xC0: case 0xC1: case 0xC2: read_sof(marker); break;
                case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                    throw JpegError{"unsupported JPEG: lossless, hierarchical or arithmetic-coded frame"};
                case 0xDD: restart_interval = u16(); break;
                case 0xE0: if (len >= 7 && std::memcmp(data + pos, "JFIF", 5) == 0) jfif = true; break;
                case 0xEE:
                    if (len >= 14 && std::memcmp(data + pos, "Adobe", 5) == 0) { adobe = true; adobe_transform = data[pos + 11]; }
                    break;
                case 0xDA:
                    pos = seg + 2;
                    read_scan();
                    any_scan = true;
                    continue;  // read_scan left pos at the next marker
                default: break;
            }
            pos = end;
        }
        if (!have_frame || !any_scan) throw JpegError{"corrupt JPEG: no image data"};
        reconstruct_planes();
        to_rgb(rgb);
        w = (uint32_t)width; h = (uint32_t)height;
    }
};

// ================================================================ PNG (RFC 2083) + zlib inflate (RFC 1950 / 1951) ===============

struct PngError {
    std::string message;
};

struct Inflater {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc = 0;
    int count = 0;
    std::vector<uint8_t>& out;

    void need(int nbits) {
        while (count < nbits) {
            if (p >= end) throw PngError{"truncated deflate stream"};
            acc |= (uint64_t)*p++ << count;
            count += 8;
        }
    }
    uint32_t bits(int nbits) {
        if (nbits == 0) return 0;
        need(nbits);
        const uint32_t v = (uint32_t)(acc & ((1ull << nbits) - 1));
        acc >>= nbits; count -= nbits;
        return v;
    }
    struct Code {
        uint16_t count[16] = {};
        uint16_t symbol[288] = {};
        void build(const uint8_t* lengths, int n) {
            std::fill(count, count + 16, (uint16_t)0);
            for (int i = 0; i < n; ++i) count[lengths[i]]++;
            count[0] = 0;
            uint16_t offs[16] = {};
            for (int len = 1; len < 16; ++len) offs[len] = (uint16_t)(offs[len - 1] + count[len - 1]);
            for (int i = 0; i < n; ++i)
                if (lengths[i]) symbol[offs[lengths[i]]++] = (uint16_t)i;
        }
    };
    int decode(const Code& c) {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len < 16; ++len) {
            code |= (int)bits(1);
            const int cnt = c.count[len];
            if (code - cnt < first) return c.symbol[index + (code - first)];
            index += cnt; first += cnt;
            first <<= 1; code <<= 1;
        }
        throw PngError{"corrupt deflate stream: bad code"};
    }
    void block(const Code& lit, const Code& dist) {
        static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        static const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        for (;;) {
            int sym = decode(lit);
            if (sym < 256) { out.push_back((uint8_t)sym); continue; }
            if (sym == 256) return;
            sym -= 257;
            if (sym >= 29) throw PngError{"corrupt deflate stream: bad length symbol"};
            const int len = lbase[sym] + (int)bits(lext[sym]);
            const int ds = decode(dist);
            if (ds >= 30) throw PngError{"corrupt deflate stream: bad distance symbol"};
            const size_t d = dbase[ds] + bits(dext[ds]);
            if (d > out.size()) throw PngError{"corrupt deflate stream: distance too far back"};
            const size_t from = out.size() - d;
            for (int i = 0; i < len; ++i) out.push_back(out[from + i]);
        }
    }
    void run() {
        bits(16);  // zlib header (CMF, FLG); the window size does not matter to a decoder with the whole output in memory
        int last;
        do {
            last = (int)bits(1);
            const int type = (int)bits(2);
            if (type == 0) {
                acc = 0; count = 0;
                if (end - p < 4) throw PngError{"truncated deflate stream"};
                const size_t len = p[0] | p[1] << 8;
                p += 4;
                if ((size_t)(end - p) < len) throw PngError{"truncated deflate stream"};
                out.insert(out.end(), p, p + len);
                p += len;
            } else if (type == 1) {
                uint8_t l[288];
                for (int i = 0; i < 144; ++i) l[i] = 8;
                for (int i = 144; i < 256; ++i) l[i] = 9;
                for (int i = 256; i < 280; ++i) l[i] = 7;
                for (int i = 280; i < 288; ++i) l[i] = 8;
                Code lit, dist;
                lit.build(l, 288);
                uint8_t d[30];
                std::fill(d, d + 30, (uint8_t)5);
                dist.build(d, 30);
                block(lit, dist);
            } else if (type == 2) {
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                const int nlen = (int)bits(5) + 257, ndist = (int)bits(5) + 1, ncode = (int)bits(4) + 4;
                if (nlen > 286 || ndist > 30) throw PngError{"corrupt deflate stream: bad code counts"};
                uint8_t l[320] = {};
                for (int i = 0; i < ncode; ++i) l[order[i]] = (uint8_t)bits(3);
                Code lencode;
                lencode.build(l, 19);
                std::fill(l, l + 320, (uint8_t)0);
                for (int i = 0; i < nlen + ndist;) {
                    int sym = decode(lencode);
                    if (sym < 16) { l[i++] = (uint8_t)sym; continue; }
                    int prev = 0, rep;
                    if (sym == 16) {
                        if (i == 0) throw PngError{"corrupt deflate stream: repeat without a first length"};
                        prev = l[i - 1];
                        rep = 3 + (int)bits(2);
                    } else if (sym == 17) rep = 3 + (int)bits(3);
                    else rep = 11 + (int)bits(7);
                    if (i + rep > nlen + ndist) throw PngError{"corrupt deflate stream: too many lengths"};
                    while (rep--) l[i++] = (uint8_t)prev;
                }
                Code lit, dist;
                lit.build(l, nlen);
                dist.build(l + nlen, ndist);
                block(lit, dist);
            } else {
                throw PngError{"corrupt deflate stream: bad block type"};
            }
        } while (!last);
    }
};

inline uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

void png_run(const uint8_t* data, size_t n, uint32_t& w, uint32_t& h, std::vector<uint8_t>& rgb) {
    static const uint8_t magic[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (n < 8 || std::memcmp(data, magic, 8) != 0) throw PngError{"not a PNG file"};
    size_t pos = 8;
    int depth = 0, ctype = 0, interlace = 0;
    bool have_header = false;
    std::vector<uint8_t> idat, palette;
    while (pos + 8 <= n) {
        const uint32_t len = be32(data + pos);
        const uint8_t* type = data + pos + 4;
        const uint8_t* body = data + pos + 8;
        if (pos + 12 + (size_t)len > n) throw PngError{"truncated PNG"};
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len < 13) throw PngError{"corrupt PNG: short IHDR"};
            w = be32(body); h = be32(body + 4);
            depth = body[8]; ctype = body[9]; interlace = body[12];
            have_header = true;
        } else if (!std::memcmp(type, "PLTE", 4)) {
            palette.assign(body, body + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_header || w == 0 || h == 0) throw PngError{"corrupt PNG: no header"};
    if (interlace) throw PngError{"unsupported PNG: Adam7 interlacing"};
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: throw PngError{"corrupt PNG: bad colour type"};
    }
    if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4))) || (ctype == 3 && depth == 16))
        throw PngError{"corrupt PNG: bad bit depth"};
    const size_t bpp = std::max<size_t>(1, (size_t)channels * depth / 8);            // filter unit
    const size_t row_bytes = ((size_t)w * channels * depth + 7) / 8;
    std::vector<uint8_t> raw;
    raw.reserve((row_bytes + 1) * h);
    Inflater inf{idat.data(), idat.data() + idat.size(), 0, 0, raw};
    inf.run();
    if (raw.size() < (row_bytes + 1) * h) throw PngError{"truncated PNG: image data too short"};
    // undo the scan-line filters in place (RFC 2083 section 6)
    std::vector<uint8_t> zero(row_bytes, 0);
    for (uint32_t y = 0; y < h; ++y) {
        uint8_t* cur = raw.data() + (size_t)y * (row_bytes + 1) + 1;
        const uint8_t* up = y ? cur - (row_bytes + 1) : zero.data();
        const int filter = cur[-1];
        for (size_t i = 0; i < row_bytes; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
            int pred = 0;
            switch (filter) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: {
                    const int pp = a + b - c, pa = std::abs(pp - a), pb = std::abs(pp - b), pc = std::abs(pp - c);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: throw PngError{"corrupt PNG: bad filter type"};
            }
            cur[i] = (uint8_t)(cur[i] + pred);
        }
    }
    // to RGB8 (DynamicImage::to_rgb8: alpha dropped, grey replicated, 16-bit samples rounded to 8)
    rgb.resize((size_t)w * h * 3);
    auto sample = [&](const uint8_t* row, size_t index) -> uint32_t {  // channel sample `index` of the row, as stored
        if (depth == 8) return row[index];
        if (depth == 16) return (uint32_t)row[2 * index] << 8 | row[2 * index + 1];
        const size_t bit = index * depth;
        return (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
    };
    auto to8 = [&](uint32_t v) -> uint8_t {
        if (depth == 8) return (uint8_t)v;
        if (depth == 16) return (uint8_t)((v + 128) / 257);
        return (uint8_t)(v * 255 / ((1u << depth) - 1u));
    };
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t* row = raw.data() + (size_t)y * (row_bytes + 1) + 1;
        uint8_t* px = &rgb[(size_t)y * w * 3];
        for (uint32_t x = 0; x < w; ++x, px += 3) {
            if (ctype == 3) {
                const uint32_t idx = sample(row, x);
                if ((size_t)idx * 3 + 2 >= palette.size()) throw PngError{"corrupt PNG: palette index out of range"};
                px[0] = palette[idx * 3]; px[1] = palette[idx * 3 + 1]; px[2] = palette[idx * 3 + 2];
            } else if (ctype == 0 || ctype == 4) {
                px[0] = px[1] = px[2] = to8(sample(row, (size_t)x * channels));
            } else {
                px[0] = to8(sample(row, (size_t)x * channels));
                px[1] = to8(sample(row, (size_t)x * channels + 1));
                px[2] = to8(sample(row, (size_t)x * channels + 2));
            }
        }
    }
}

}  // namespace

bool decode_jpeg(const uint8_t* data, size_t n, uint32_t& width, uint32_t& height, std::vector<uint8_t>& rgb, std::string& err) {
    try {
        Decoder d(data, n);
        d.run(width, height, rgb);
        return true;
    } catch (const JpegError& e) {
        err = e.message;
        return false;
    }
}

bool decode_png(const uint8_t* data, size_t n, uint32_t& width, uint32_t& height, std::vector<uint8_t>& rgb, std::string& err) {
    try {
        png_run(data, n, width, height, rgb);
        return true;
    } catch (const PngError& e) {
        err = e.message;
        return false;
    }
}

bool decode_image(const uint8_t* data, size_t n, uint32_t& width, uint32_t& height, std::vector<uint8_t>& rgb, std::string& err) {
    if (n >= 2 && data[0] == 0xFF && data[1] == 0xD8) return decode_jpeg(data, n, width, height, rgb, err);
    if (n >= 4 && data[0] == 0x89 && data[1] == 'P' && data[2] == 'N' && data[3] == 'G') return decode_png(data, n, width, height, rgb, err);
    err = "unsupported image format (built in: JPEG, PNG, binary PPM)";
    return false;
}

}  // namespace cray
