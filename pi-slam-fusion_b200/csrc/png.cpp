// png.cpp — minimal PNG encoder for Map2D::save (the reference calls cv::imwrite, Map2DCPU.cpp:562,
// MultiBandMap2DCPU.cpp:841).  8-bit BGR/BGRA in, RGB/RGBA PNG out, zlib deflate, filter type 0.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <vector>

static void put32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x);
}
static bool write_chunk(FILE* f, const char* tag, const uint8_t* data, size_t n) {
    std::vector<uint8_t> hdr;
    put32(hdr, (uint32_t)n);
    uint32_t crc = crc32(0L, (const Bytef*)tag, 4);
    if (n) crc = crc32(crc, data, (uInt)n);
    std::vector<uint8_t> tail;
    put32(tail, crc);
    return fwrite(hdr.data(), 1, 4, f) == 4 && fwrite(tag, 1, 4, f) == 4 && (n == 0 || fwrite(data, 1, n, f) == n) &&
           fwrite(tail.data(), 1, 4, f) == 4;
}

int m2d_write_png(const char* path, const uint8_t* px, int w, int h, int channels) {
    if (channels != 3 && channels != 4) return -1;
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    bool ok = fwrite(sig, 1, 8, f) == 8;
    std::vector<uint8_t> ihdr;
    put32(ihdr, (uint32_t)w); put32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(channels == 4 ? 6 : 2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    ok = ok && write_chunk(f, "IHDR", ihdr.data(), ihdr.size());
    // stream rows through deflate in bands so multi-gigapixel mosaics do not need a second full copy
    z_stream zs{};
    if (deflateInit(&zs, 3) != Z_OK) { fclose(f); return -1; }
    const size_t row = (size_t)w * channels + 1;
    std::vector<uint8_t> in(row), outbuf(1 << 20);
    for (int y = 0; y < h && ok; y++) {
        in[0] = 0;
        const uint8_t* s = px + (size_t)y * w * channels;
        for (int x = 0; x < w; x++) {
            in[1 + x * channels + 0] = s[x * channels + 2];
            in[1 + x * channels + 1] = s[x * channels + 1];
            in[1 + x * channels + 2] = s[x * channels + 0];
            if (channels == 4) in[1 + x * 4 + 3] = s[x * 4 + 3];
        }
        zs.next_in = in.data();
        zs.avail_in = (uInt)row;
        int flush = (y == h - 1) ? Z_FINISH : Z_NO_FLUSH;
        do {
            zs.next_out = outbuf.data();
            zs.avail_out = (uInt)outbuf.size();
            int r = deflate(&zs, flush);
            if (r == Z_STREAM_ERROR) { ok = false; break; }
            size_t have = outbuf.size() - zs.avail_out;
            if (have) ok = ok && write_chunk(f, "IDAT", outbuf.data(), have);
        } while (zs.avail_out == 0 && ok);
    }
    deflateEnd(&zs);
    ok = ok && write_chunk(f, "IEND", nullptr, 0);
    ok = (fclose(f) == 0) && ok;
    return ok ? 0 : -1;
}
