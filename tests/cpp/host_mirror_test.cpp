// C++ host-mirror test: the reference's own test expectations, restated against include/melspec_b200.hpp (the compiled-
// language host side above the C ABI).  Usage: host_mirror_test <tests/golden dir> [nogpu]
//   nogpu : only the device-free checks (sizes, error mapping when no CUDA device / bad config)
//   else  : src/rb.rs:134-179 (stream vs rust_jfk_golden.npy), src/cuda.rs:488-545 (test signal, shapes),
//           tests/readme_examples.rs:11-99 (shape contracts), src/fbank.rs:439-535, src/mel.rs:943-961,
//           src/vad.rs:621-668 (known answers), quantized_mel_golden.tga
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "melspec_b200.hpp"

using namespace mel_spec;

static int g_fail = 0;
#define CHECK(cond, msg)                                                                   \
    do {                                                                                   \
        if (!(cond)) { std::printf("FAIL %s:%d  %s  [%s]\n", __FILE__, __LINE__, #cond, std::string(msg).c_str()); ++g_fail; } \
    } while (0)

static std::vector<uint8_t> read_file(const std::string& p) {
    std::ifstream f(p, std::ios::binary);
    if (!f) { std::printf("cannot open %s\n", p.c_str()); std::exit(2); }
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

// minimal .npy reader: little-endian f32, C order
static std::vector<float> read_npy_f32(const std::string& p, std::vector<size_t>* shape = nullptr, bool* fortran = nullptr) {
    std::vector<uint8_t> b = read_file(p);
    if (b.size() < 10 || std::memcmp(b.data(), "\x93NUMPY", 6) != 0) { std::printf("%s: not an npy file\n", p.c_str()); std::exit(2); }
    const size_t hlen = b[8] | (b[9] << 8), off = 10 + hlen;
    std::string hdr(b.begin() + 10, b.begin() + off);
    if (hdr.find("'<f4'") == std::string::npos) { std::printf("%s: unexpected dtype\n", p.c_str()); std::exit(2); }
    const bool f_order = hdr.find("'fortran_order': True") != std::string::npos;
    if (fortran) *fortran = f_order;
    else if (f_order) { std::printf("%s: unexpected Fortran order\n", p.c_str()); std::exit(2); }
    if (shape) {
        shape->clear();
        const size_t a = hdr.find("'shape': (") + 10, e = hdr.find(')', a);
        std::stringstream ss(hdr.substr(a, e - a));
        std::string tok;
        while (std::getline(ss, tok, ',')) if (tok.find_first_of("0123456789") != std::string::npos) shape->push_back(std::stoul(tok));
    }
    std::vector<float> v((b.size() - off) / 4);
    std::memcpy(v.data(), b.data() + off, v.size() * 4);
    return v;
}

int main(int argc, char** argv) {
    if (argc < 2) { std::printf("usage: %s <golden dir> [nogpu]\n", argv[0]); return 2; }
    const std::string g = argv[1];
    const bool nogpu = argc > 2 && std::string(argv[2]) == "nogpu";

    // ---- device-free checks
    CHECK(melspec_interleaved_width(1097, 0) == 1097 && melspec_interleaved_width(1097, 2) == 1098 && melspec_interleaved_width(5, 3) == -1, "");
    CHECK(melspec_tga_size(80, 1100) == 88026 && melspec_tga_size(80, 65536) == -1 && melspec_tga_size(80, 65535) > 0, "");
    {
        bool thrown = false;
        try { CudaMelSpectrogram bad(0, 160, 16000.0, 80); } catch (const CudaError& e) { thrown = e.kind == CudaError::Kind::Unavailable; }
        CHECK(thrown, "zero sizes => CudaError::Unavailable (src/cuda.rs:45-49)");
    }
    EdgeInfo e1; e1.intersected_columns = {5};
    EdgeInfo e2; e2.intersected_columns = {5, 6};
    CHECK(!vad_on(e1, 1) && vad_on(e2, 1) && !vad_on(EdgeInfo(), 1), "vad_on quirk (src/vad.rs:226-249)");
    if (nogpu) {
        bool unavailable = false;
        try { CudaMelSpectrogram m(400, 160, 16000.0, 80); } catch (const CudaError& e) { unavailable = e.kind == CudaError::Kind::Unavailable; }
        std::printf("nogpu: construction %s\n", unavailable ? "throws CudaError::Unavailable (no device)" : "succeeded (a device is present)");
        std::printf(g_fail ? "FAILED (%d)\n" : "ok\n", g_fail);
        return g_fail ? 1 : 0;
    }

    const std::vector<float> jfk = read_npy_f32(g + "/jfk_pcm_f32.npy");
    CHECK(jfk.size() == 176000, "");

    // ---- CudaMelSpectrogram: shapes, silence, empty input (tests/readme_examples.rs, src/cuda.rs:91-93)
    CudaMelSpectrogram mel(400, 160, 16000.0, 80);
    {
        auto fr = mel.compute_mel_spectrogram(jfk);
        CHECK(fr.size() == 1098 && fr[0].size() == 80, "frames x mels");
        auto z = mel.compute_mel_spectrogram(std::vector<float>(16000, 0.0f));
        CHECK(z.size() == 98, "");
        bool all = true;
        for (auto& r : z) for (float v : r) all = all && v == -1.5f;
        CHECK(all, "silence => -1.5 (1e-10 floor)");
        CHECK(mel.compute_mel_spectrogram(std::vector<float>(399)).empty() && mel.compute_mel_spectrogram({}).empty(), "short input => no frames");
        CHECK(mel.max_frames_per_batch() == 8192, "src/cuda.rs:150-155");
        // interleave == transpose of the frame-major result
        size_t w = 0;
        std::vector<uint8_t> tga;
        auto img = mel.interleave_frames(jfk, 1100, &tga, &w);
        CHECK(w == 1100 && img.size() == 80 * 1100 && tga.size() == 88026, "");
        float md = 0;
        for (size_t k = 0; k < 1098; ++k) for (size_t m = 0; m < 80; ++m) md = std::fmax(md, std::fabs(img[m * 1100 + k] - fr[k][m]));
        CHECK(md == 0.0f, "interleaved image == frames transposed");
        bool pad0 = true;
        for (size_t m = 0; m < 80; ++m) pad0 = pad0 && img[m * 1100 + 1098] == 0.0f && img[m * 1100 + 1099] == 0.0f;
        CHECK(pad0, "zero padding columns (src/mel.rs:506-516)");
        CHECK(mel.parse_tga_8bit(tga).size() == 88000, "");
    }

    // ---- stream path vs rust_jfk_golden.npy (src/rb.rs:134-179): fft 512, hop 160, 80 mels, (80, 1097)
    {
        std::vector<size_t> shp;
        const std::vector<float> gold = read_npy_f32(g + "/rust_jfk_golden.npy", &shp);
        CHECK(shp.size() == 2 && shp[0] == 80 && shp[1] == 1097, "");
        RingBuffer rb(MelConfig(512, 160, 80, 16000.0), 1 << 20);
        size_t k = 0;
        float md = 0;
        for (size_t off = 0; off < jfk.size(); off += 32) {           // 128-byte reads like the reference's test
            rb.add_frame(std::vector<float>(jfk.begin() + off, jfk.begin() + std::min(off + 32, jfk.size())));
            while (auto f = rb.maybe_mel()) {
                if (k < 1097) for (size_t m = 0; m < 80; ++m) md = std::fmax(md, std::fabs((*f)[m] - gold[m * 1097 + k]));
                ++k;
            }
        }
        CHECK(k == 1097, "stream frame count " + std::to_string(k));
        CHECK(md <= 1e-4f, "max|stream - golden| = " + std::to_string(md));
        std::printf("ringbuffer fft512 vs rust_jfk_golden.npy: %zu frames, max abs diff %.3g\n", k, md);
    }

    // ---- Spectrogram::add's own contract (src/stft.rs:175-194: fft 8, hop 4 -> None, None (7 < 8 true samples), Some), and fed
    //      whole hops it equals the batch path on samples[80..] (stream offset c = 80 for 400 / 160)
    {
        Spectrogram sp(8, 4, 4, 16000.0);
        MelSpectrogram ms(8, 16000.0, 4);
        CHECK(!sp.add({1.f, 2.f, 3.f}).has_value(), "3 samples: None");
        CHECK(!sp.add({1.f, 2.f, 3.f, 4.f}).has_value(), "7 samples: None");
        auto f = sp.add({1.f, 2.f, 3.f, 4.f});
        CHECK(f.has_value() && ms.add(*f).size() == 4, "11 samples: Some");
        bool threw = false;
        try { sp.add(std::vector<float>(5, 0.f)); } catch (const std::invalid_argument&) { threw = true; }
        CHECK(threw, "frames must be <= hop_size");
        Spectrogram s400(400, 160);
        MelSpectrogram m400(400, 16000.0, 80);
        std::vector<float> tail(jfk.begin() + 80, jfk.begin() + 80 + 160 * 60 + 240);
        auto batch = mel.compute_mel_spectrogram(tail);
        size_t k = 0;
        float md = 0;
        for (size_t off = 0; off + 160 <= 160 * 62; off += 160) {
            auto fr = s400.add(std::vector<float>(jfk.begin() + off, jfk.begin() + off + 160));
            if (!fr) continue;
            auto col = m400.add(*fr);
            if (k < batch.size()) for (size_t m = 0; m < 80; ++m) md = std::fmax(md, std::fabs(col[m] - batch[k][m]));
            ++k;
        }
        CHECK(k == 60 && md <= 5e-5f, "Spectrogram::add fed whole hops == batch path on samples[80..]: " + std::to_string(k) + " frames, " + std::to_string(md));
        std::vector<int16_t> pcm16(16000);
        std::vector<float> pcmf(16000);
        for (size_t i = 0; i < pcm16.size(); ++i) { pcm16[i] = (int16_t)std::lround(jfk[i] * 32767.0f); pcmf[i] = (float)pcm16[i] / 32768.0f; }
        CHECK(mel.compute_mel_spectrogram_i16(pcm16) == mel.compute_mel_spectrogram(pcmf), "int16 entry == f32 entry on x / 32768");
    }

    // ---- PCM -> TGA vs quantized_mel_golden.tga (fft 400 stream framing = batch framing on samples[80..])
    {
        const std::vector<uint8_t> gold = read_file(g + "/quantized_mel_golden.tga");
        std::vector<float> tail(jfk.begin() + 80, jfk.end());
        std::vector<uint8_t> tga;
        size_t w = 0;
        mel.interleave_frames(tail, 0, &tga, &w);
        CHECK(w == 1098 && tga.size() == 26 + 80 * 1098, "");
        size_t diff = 0; int mx = 0;
        for (size_t m = 0; m < 80; ++m)
            for (size_t k = 0; k < 1098; ++k) {
                const int d = std::abs((int)tga[26 + m * 1098 + k] - (int)gold[26 + m * 1100 + k + 2]);
                diff += d != 0; mx = std::max(mx, d);
            }
        CHECK(mx <= 1 && diff <= 878, "levels differing: " + std::to_string(diff) + ", max " + std::to_string(mx));
        std::printf("mel_tga vs quantized_mel_golden.tga: %zu of 87840 pixels differ (by at most %d level)\n", diff, mx);
        // VAD on the golden image with the reference's settings for it (src/vad.rs:701-716): runs and partitions all columns
        auto img = mel.parse_tga_8bit(gold);
        DetectionSettings s; s.min_energy = 1.0; s.min_y = 3; s.min_x = 6; s.min_mel = 0;
        auto e = mel.vad_boundaries(img, 80, s);
        CHECK(e.intersected().size() + e.non_intersected().size() == 1098 && !e.intersected().empty(), "");
        auto acts = mel.vad_activities(img, 80, DetectionSettings());
        CHECK(acts.size() == 1100 - 5 + 1 && acts[0].frame_index == 4 && acts[0].window_columns == 3, "");
    }

    // ---- VAD known answers (src/vad.rs:621-668)
    {
        DetectionSettings s; s.min_energy = 1.0; s.min_y = 10; s.min_x = 10; s.min_mel = 0;
        const char* blank[] = {"21168", "23760", "41492", "41902", "63655", "7497", "39744"};
        const char* speech[] = {"11648", "2889", "4694", "4901", "27125"};
        for (const char* id : blank) {
            auto img = mel.parse_tga_8bit(read_file(g + "/vad/blank/frame_" + id + ".tga"));
            CHECK(!vad_on(mel.vad_boundaries(img, 80, s), 10), std::string("blank ") + id);
        }
        for (const char* id : speech) {
            auto img = mel.parse_tga_8bit(read_file(g + "/vad/speech/frame_" + id + ".tga"));
            CHECK(vad_on(mel.vad_boundaries(img, 80, s), 10), std::string("speech ") + id);
        }
    }

    // ---- Fbank (src/fbank.rs:439-535: shape, finiteness, variance; distance to the knf golden as in SURVEY 8c)
    {
        Fbank fb;
        size_t t = 0;
        auto f = fb.compute(jfk, &t);
        CHECK(t == 1098 && f.size() == 1098 * 80, "");
        std::vector<size_t> shp;
        bool f_order = false;
        const std::vector<float> gold = read_npy_f32(g + "/kaldi_fbank_jfk.npy", &shp, &f_order);   // (80, 1098), either order
        double mean = 0, var = 0, md = 0;
        bool finite = true;
        for (float v : f) { mean += v; finite = finite && std::isfinite(v); }
        mean /= f.size();
        for (float v : f) var += (v - mean) * (v - mean);
        var /= f.size();
        for (size_t k = 0; k < 1098; ++k) for (size_t m = 0; m < 80; ++m) md = std::fmax(md, std::fabs(f[k * 80 + m] - gold[f_order ? k * 80 + m : m * 1098 + k]));
        CHECK(finite && var > 0.1 && md < 2.5e-2, "var " + std::to_string(var) + " max diff " + std::to_string(md));
        std::printf("fbank vs kaldi_native_fbank golden: max abs diff %.3g (the reference itself: 1.5e-2)\n", md);
    }

    // ---- BatchLogMelSpectrogram shape contract (src/mel.rs:943-961)
    {
        BatchLogMelConfig c; c.n_mels = 128;
        BatchLogMelSpectrogram b(c);
        size_t rows = 0, cols = 0;
        auto f = b.compute_flat(std::vector<float>(16000, 0.01f), &rows, &cols);
        CHECK(rows == 128 && cols == 101 && f.size() == 128 * 101, "");
    }

    // ---- Spectrogram::compute_mel_spectrogram (src/stft.rs:119-138) and the reference's CUDA test signal (src/cuda.rs:494-502)
    {
        std::vector<float> x(16000);
        for (size_t i = 0; i < x.size(); ++i) {
            const double t = (double)i / 16000.0;
            x[i] = (float)(0.6 * std::sin(2 * M_PI * 220 * t) + 0.25 * std::sin(2 * M_PI * 440 * t) + 0.10 * std::sin(2 * M_PI * 880 * t) +
                           0.05 * std::sin(2 * M_PI * 1760 * t));
        }
        auto fr = Spectrogram::compute_mel_spectrogram(x, 400, 160, 80, 16000.0);
        CHECK(fr.size() == 98 && fr[0].size() == 80, "");
    }

    std::printf(g_fail ? "FAILED (%d)\n" : "ok\n", g_fail);
    return g_fail ? 1 : 0;
}
