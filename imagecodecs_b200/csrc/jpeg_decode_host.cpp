// jpeg_decode_host.cpp -- host half of the decoder: the marker loop of njDecode (jpeg_dec.h:880-908)
// with its segment parsers, written against a bounds-checked cursor, and the split of the scan
// into restart intervals.  No pixels are touched here.
#include "jpeg_decode.h"

#include <string.h>

namespace jd {
namespace {

struct Cursor {
    const uint8_t* p;
    size_t left;
    int error = kOk;
    bool skip(size_t n) { if (n > left) { error = kSyntaxError; return false; } p += n; left -= n; return true; }
    int be16(size_t at) const { return (p[at] << 8) | p[at + 1]; }
};

// njDecodeLength (:511-518): the segment length, checked against the rest of the file; afterwards `len`
// bytes of payload follow the cursor
bool segment(Cursor& c, size_t* len)
{
    if (c.left < 2) { c.error = kSyntaxError; return false; }
    const size_t l = (size_t)c.be16(0);
    if (l > c.left || l < 2) { c.error = kSyntaxError; return false; }
    c.skip(2);
    *len = l - 2;
    return true;
}

int parse_sof(Cursor& c, Info* I)      // njDecodeSOF (:520-571)
{
    size_t len;
    if (!segment(c, &len)) return c.error;
    if (len < 6) return kSyntaxError;
    if (c.p[0] != 8) return kUnsupported;
    I->height = c.be16(1);
    I->width = c.be16(3);
    if (!I->width || !I->height) return kSyntaxError;
    I->ncomp = c.p[5];
    if (I->ncomp != 1 && I->ncomp != 3) return kUnsupported;
    if (len < 6 + (size_t)I->ncomp * 3) return kSyntaxError;
    int ssxmax = 0, ssymax = 0;
    for (int i = 0; i < I->ncomp; ++i) {
        const uint8_t* q = c.p + 6 + 3 * i;
        Component& k = I->comp[i];
        k.cid = q[0];
        if (!(k.ssx = q[1] >> 4)) return kSyntaxError;
        if (k.ssx & (k.ssx - 1)) return kUnsupported;           // non-power of two
        if (!(k.ssy = q[1] & 15)) return kSyntaxError;
        if (k.ssy & (k.ssy - 1)) return kUnsupported;
        if ((k.qtsel = q[2]) & 0xFC) return kSyntaxError;
        I->qtused |= 1 << k.qtsel;
        if (k.ssx > ssxmax) ssxmax = k.ssx;
        if (k.ssy > ssymax) ssymax = k.ssy;
    }
    if (I->ncomp == 1) I->comp[0].ssx = I->comp[0].ssy = ssxmax = ssymax = 1;
    I->mbsizex = ssxmax << 3;
    I->mbsizey = ssymax << 3;
    I->mbwidth = (I->width + I->mbsizex - 1) / I->mbsizex;
    I->mbheight = (I->height + I->mbsizey - 1) / I->mbsizey;
    size_t blocks = 0, bytes = 0;
    for (int i = 0; i < I->ncomp; ++i) {
        Component& k = I->comp[i];
        k.width = (I->width * k.ssx + ssxmax - 1) / ssxmax;
        k.height = (I->height * k.ssy + ssymax - 1) / ssymax;
        k.stride = I->mbwidth * k.ssx << 3;
        if ((k.width < 3 && k.ssx != ssxmax) || (k.height < 3 && k.ssy != ssymax)) return kUnsupported;
        k.bw = I->mbwidth * k.ssx;
        k.bh = I->mbheight * k.ssy;
        k.coef_off = blocks;
        k.plane_off = bytes;
        blocks += (size_t)k.bw * k.bh;
        bytes += (size_t)k.stride * ((size_t)k.bh << 3);
    }
    I->n_blocks = blocks;
    I->plane_bytes = bytes;
    I->n_mcus = I->mbwidth * I->mbheight;
    c.skip(len);
    return kOk;
}

// njDecodeDHT (:573-614) over one segment payload [p, p + len): fills the 65536-entry tables
int dht_payload(const uint8_t* p, size_t len, std::vector<uint16_t>* out)
{
    if (out->empty()) out->assign(4 * 65536, 0);
    while (len >= 17) {
        int i = p[0];
        if (i & 0xEC) return kSyntaxError;
        if (i & 0x02) return kUnsupported;
        i = (i | (i >> 3)) & 3;                        // combined DC/AC + table id
        uint8_t counts[16];
        memcpy(counts, p + 1, 16);
        p += 17; len -= 17;
        uint16_t* vlc = &(*out)[(size_t)i * 65536];
        int remain = 65536, spread = 65536;
        for (int codelen = 1; codelen <= 16; ++codelen) {
            spread >>= 1;
            const int currcnt = counts[codelen - 1];
            if (!currcnt) continue;
            if (len < (size_t)currcnt) return kSyntaxError;
            remain -= currcnt << (16 - codelen);
            if (remain < 0) return kSyntaxError;
            for (int k = 0; k < currcnt; ++k) {
                const uint16_t e = (uint16_t)((codelen << 8) | p[k]);
                for (int j = spread; j; --j) *vlc++ = e;
            }
            p += currcnt; len -= (size_t)currcnt;
        }
        while (remain--) *vlc++ = 0;
    }
    return len ? kSyntaxError : kOk;
}

// a DHT segment: remembered verbatim (length-prefixed); the tables are built from these bytes
int parse_dht(Cursor& c, Info* I)
{
    size_t len;
    if (!segment(c, &len)) return c.error;
    I->dht.push_back((uint8_t)(len >> 8)); I->dht.push_back((uint8_t)len);
    I->dht.insert(I->dht.end(), c.p, c.p + len);
    c.skip(len);
    return kOk;
}

int parse_dqt(Cursor& c, Info* I)      // njDecodeDQT (:616-631)
{
    size_t len;
    if (!segment(c, &len)) return c.error;
    while (len >= 65) {
        const int i = c.p[0];
        if (i & 0xFC) return kSyntaxError;
        I->qtavail |= 1 << i;
        memcpy(I->qtab[i], c.p + 1, 64);
        c.skip(65); len -= 65;
    }
    return len ? kSyntaxError : kOk;
}

int parse_sos(Cursor& c, Info* I)      // header part of njDecodeScan (:674-692)
{
    size_t len;
    if (!segment(c, &len)) return c.error;
    if (len < (size_t)(1 + 2 * I->ncomp + 3)) return kSyntaxError;
    if (c.p[0] != I->ncomp) return kUnsupported;
    for (int i = 0; i < I->ncomp; ++i) {
        const uint8_t* q = c.p + 1 + 2 * i;
        if (q[0] != I->comp[i].cid) return kSyntaxError;
        if (q[1] & 0xEE) return kSyntaxError;
        I->comp[i].dctabsel = q[1] >> 4;
        I->comp[i].actabsel = (q[1] & 1) | 2;
    }
    const uint8_t* t = c.p + 1 + 2 * I->ncomp;
    if (t[0] || t[1] != 63 || t[2]) return kUnsupported;
    c.skip(len);
    return kOk;
}

}  // namespace

int build_vlc_tables(const std::vector<uint8_t>& dht, std::vector<uint16_t>* vlc)
{
    vlc->clear();
    for (size_t at = 0; at + 2 <= dht.size();) {
        const size_t len = ((size_t)dht[at] << 8) | dht[at + 1];
        const int rc = dht_payload(dht.data() + at + 2, len, vlc);
        if (rc != kOk) return rc;
        at += 2 + len;
    }
    return vlc->empty() ? (int)kSyntaxError : (int)kOk;
}

int mcu_block_map(const Info& I, McuBlock* map)
{
    int n = 0;
    for (int c = 0; c < I.ncomp; ++c) {
        const Component& k = I.comp[c];
        for (int sby = 0; sby < k.ssy; ++sby)
            for (int sbx = 0; sbx < k.ssx; ++sbx) {
                if (n == kMaxBlocksPerMcu) return 0;
                map[n++] = {(unsigned long long)k.coef_off + (unsigned long long)sby * k.bw + sbx, k.ssy * k.bw, k.ssx, c, k.dctabsel, k.actabsel, 0};
            }
    }
    return n;
}

int subsequence_log2(const Info& I, size_t batch_scan_bytes)
{
    const size_t bytes = I.scan_end - I.scan_off;
    if (I.rstinterval || !I.clean_stuffing || bytes < 512 || bytes >= ((size_t)1 << 28)) return 0;
    McuBlock map[kMaxBlocksPerMcu];
    if (!mcu_block_map(I, map)) return 0;
    // enough subsequences to fill the GPU when the call is small, 128-byte ones (the size the literature settles on) when it is large
    return batch_scan_bytes >= ((size_t)8 << 20) ? 7 : batch_scan_bytes >= ((size_t)2 << 20) ? 6 : 5;
}

int parse(const uint8_t* jpeg, size_t size, Info* I, bool build_vlc, bool header_only)
{
    size &= 0x7FFFFFFF;
    if (size < 2 || jpeg[0] != 0xFF || jpeg[1] != 0xD8) return kNoJpeg;
    Cursor c{jpeg + 2, size - 2};
    bool have_sof = false;
    for (;;) {
        if (c.left < 2 || c.p[0] != 0xFF) return kSyntaxError;
        const int marker = c.p[1];
        c.skip(2);
        int rc = kOk;
        size_t len;
        switch (marker) {
            case 0xC0: rc = parse_sof(c, I); have_sof = rc == kOk; break;
            case 0xC4: rc = parse_dht(c, I); break;
            case 0xDB: rc = parse_dqt(c, I); break;
            case 0xDD:                                        // njDecodeDRI (:633-641)
                if (!segment(c, &len)) return c.error;
                if (len < 2) return kSyntaxError;
                I->rstinterval = c.be16(0);
                c.skip(len);
                break;
            case 0xDA:
                if (!have_sof || I->dht.empty()) return kSyntaxError;
                rc = parse_sos(c, I);
                if (rc != kOk) return rc;
                if (build_vlc && (rc = build_vlc_tables(I->dht, &I->vlc)) != kOk) return rc;
                goto scan;
            case 0xFE:
                if (!segment(c, &len)) return c.error;
                c.skip(len);
                break;
            default:
                if ((marker & 0xF0) != 0xE0) return kUnsupported;
                if (!segment(c, &len)) return c.error;
                c.skip(len);
        }
        if (rc != kOk) return rc;
    }
scan:
    // The entropy-coded data: split at the RSTm markers (njDecodeScan :707-715 expects RST(k mod 8) after
    // every rstinterval MCUs), ends at EOI or at the end of the file.
    I->scan_off = (size_t)(c.p - jpeg);
    if (header_only) return kOk;
    I->interval_off.clear();
    I->interval_off.push_back((uint32_t)I->scan_off);
    size_t pos = I->scan_off;
    unsigned next_rst = 0;
    size_t end = size;
    I->clean_stuffing = true;
    while (pos + 1 < size) {
        const uint8_t* f = (const uint8_t*)memchr(jpeg + pos, 0xFF, size - 1 - pos);
        if (!f) break;
        pos = (size_t)(f - jpeg);
        const int m = jpeg[pos + 1];
        if (m == 0xD9) { end = pos; break; }
        if ((m & 0xF8) == 0xD0) {
            if (!I->rstinterval || (unsigned)(m & 7) != next_rst) return kSyntaxError;
            next_rst = (next_rst + 1) & 7;
            I->interval_off.push_back((uint32_t)(pos + 2));
        } else if (m != 0x00) {
            I->clean_stuffing = false;
        }
        pos += 2;
    }
    I->scan_end = end;
    const int expected = I->rstinterval ? (I->n_mcus + I->rstinterval - 1) / I->rstinterval : 1;
    if ((int)I->interval_off.size() != expected) return kSyntaxError;
    I->interval_off.push_back((uint32_t)end);
    return kOk;
}

}  // namespace jd
