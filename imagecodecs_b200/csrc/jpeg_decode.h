// jpeg_decode.h -- what the host parser, the decode kernels and the CPU test harness share.
//
// the step on the other side of the format.  Replaces NanoJPEG as
// Image::readJpg uses it (codecs.cpp:821-849): njDecode (jpeg_dec.h:880-908) = marker loop
// (SOF0 :520-571, DHT :573-614, DQT :616-631, DRI :633-641, SOS/scan :674-718) + njConvert
// (:817-866: chroma upsampling :736-790, YCbCr -> RGB).  Everything is integer arithmetic; the
// decoded pixels are bit-identical to NanoJPEG's.
//
// Parallelism of the entropy decode comes from restart intervals where the file has them (the
// encoder's opt-in JPEG_GPU_FLAG_RESTART writes one per 24 blocks): every interval is decoded by its
// own thread.  A scan WITHOUT restart markers -- everything the reference's own encoder writes -- is
// cut into fixed-size subsequences that are decoded speculatively and brought into agreement by the
// self-synchronisation of Huffman codes (jpeg_decode.cuh, "subsequences"); only streams that are
// neither (irregular marker bytes inside the scan, exotic sampling) fall back to one thread.  The
// IDCT / upsampling / colour stages are parallel either way.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace jd {

// nj_result_t (jpeg_dec.h:118-127)
enum Result { kOk = 0, kNoJpeg = 1, kUnsupported = 2, kOutOfMem = 3, kInternalErr = 4, kSyntaxError = 5 };

struct Component {          // nj_component_t (jpeg_dec.h:296-306) + where its data lives in the work buffers
    int cid, ssx, ssy;
    int width, height, stride;      // as njDecodeSOF computes them (:558-562); stride covers whole MCUs
    int qtsel, actabsel, dctabsel;
    int bw, bh;                     // blocks per row / block rows of the padded plane
    size_t coef_off;                // first block of the component in the coefficient buffer
    size_t plane_off;               // first byte of the component's plane
};

struct Info {
    int width = 0, height = 0, ncomp = 0;
    int mbwidth = 0, mbheight = 0, mbsizex = 0, mbsizey = 0;
    int rstinterval = 0;
    Component comp[3] = {};
    uint8_t qtab[4][64] = {};
    int qtused = 0, qtavail = 0;
    std::vector<uint16_t> vlc;      // [4][65536]: bits << 8 | code (nj_vlc_code_t, :291-293), tables 0,1 = DC, 2,3 = AC
                                    // (left empty by parse(..., build_vlc = false): build_vlc_tables() makes it from `dht`)
    std::vector<uint8_t> dht;       // the payloads of the file's DHT segments, back to back: equal bytes <=> equal tables
    size_t scan_off = 0, scan_end = 0;          // entropy-coded data [scan_off, scan_end) of the file
    std::vector<uint32_t> interval_off;         // start of every restart interval (file offsets) + one past the last
    size_t n_blocks = 0, plane_bytes = 0;
    int n_mcus = 0;
    bool clean_stuffing = false;    // every FF inside the scan is followed by 00 (what every encoder writes)
};

// One block of the MCU in stream order (njDecodeScan's component / sby / sbx loops, :694-704): where its
// coefficients go and which tables decode it.
constexpr int kMaxBlocksPerMcu = 16;
struct McuBlock { unsigned long long off; int row, sx, comp, dctab, actab, pad; };
// Fills map[0..return) for the file's sampling; 0 if an MCU has more than kMaxBlocksPerMcu blocks.
int mcu_block_map(const Info& info, McuBlock* map);
// log2 of the subsequence size (bytes) for the self-synchronising decode of this file's scan, 0 = not
// eligible (restart intervals, irregular stuffing, tiny or huge scan, too many blocks per MCU).
// batch_scan_bytes: entropy-coded bytes of the whole call (more data -> longer subsequences).
int subsequence_log2(const Info& info, size_t batch_scan_bytes);

// The marker loop of njDecode up to and including the SOS header, plus the split of the scan
// into restart intervals.  Returns an nj_result_t.
int parse(const uint8_t* jpeg, size_t size, Info* info, bool build_vlc = true, bool header_only = false);   // header_only: stops at the scan (frame size, tables)
// njDecodeDHT (:573-614) over the collected DHT payloads; a batch builds each distinct table set once
int build_vlc_tables(const std::vector<uint8_t>& dht, std::vector<uint16_t>* vlc);

}  // namespace jd
