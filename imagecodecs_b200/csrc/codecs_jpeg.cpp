// codecs_jpeg.cpp -- see codecs_jpeg.h.  Host C++ in the reference's style; the only call
// into the GPU library is the drop-in twin of tje_encode_to_file.
#include "codecs_jpeg.h"

#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <fstream>
#include <vector>

#include "jpeg_gpu.h"

namespace ImageCodecs
{
	static std::string lowerExtension(const std::string& filepath)
	{
		const size_t dot = filepath.find_last_of('.');
		const size_t sep = filepath.find_last_of("/\\");
		if (dot == std::string::npos || (sep != std::string::npos && dot < sep))
			return "";
		std::string ext = filepath.substr(dot);
		for (auto& c : ext)
			c = (char)std::tolower((unsigned char)c);
		return ext;
	}

	void Image::read(std::string filepath)
	{
		// codecs.cpp:53-89, restricted to the reader the JPEG write configs need
		const std::string ext = lowerExtension(filepath);
		pendingFlip_ = pendingSwapBR_ = false;
		if (ext == ".bmp")
			readBmp(filepath, &pixels_, w_, h_, d_, type_);
		else if (ext == ".jpg" || ext == ".jpeg")
			readJpg(filepath, &pixels_, w_, h_, d_, type_);
		else
			throw std::invalid_argument("Cannot parse filetype");

		if (pixels_ == nullptr)
			throw std::runtime_error("Could not read image data");
	}

	void Image::write(std::string filepath)
	{
		// codecs.cpp:91-122: lower-cased extension picks the codec
		const std::string ext = lowerExtension(filepath);
		if (ext == ".bmp")
		{
			materialize();
			writeBmp(filepath, pixels_, w_, h_, d_, type_);
		}
		else if (ext == ".jpg" || ext == ".jpeg")
			writeJpg(filepath, pixels_, w_, h_, d_, type_);
		else
			throw std::invalid_argument("Cannot parse filetype");
	}

	// codecs.cpp:162-191: row i <- row h-1-i
	void Image::flip(unsigned char* pixels, const int w, const int h, const int d, const Type&)
	{
		const size_t lenRow = (size_t)w * d * byteSize();
		std::vector<unsigned char> row(lenRow);
		for (int i = 0; i < h / 2; ++i)
		{
			unsigned char* a = pixels + (size_t)i * lenRow;
			unsigned char* b = pixels + (size_t)(h - 1 - i) * lenRow;
			std::copy(a, a + lenRow, row.begin());
			std::copy(b, b + lenRow, a);
			std::copy(row.begin(), row.end(), b);
		}
	}

	// codecs.cpp:193-251: channels 0 and 2 of every pixel change places
	void Image::swapBR(unsigned char* pixels, const int w, const int h, const int d, const Type&)
	{
		if (d < 3)
			return;
		const size_t sz = (size_t)byteSize();
		const size_t n = (size_t)w * h;
		for (size_t p = 0; p < n; ++p)
			std::swap_ranges(pixels + (p * d) * sz, pixels + (p * d + 1) * sz, pixels + (p * d + 2) * sz);
	}

	void Image::materialize()
	{
		if (pixels_ == nullptr)
			return;
		if (pendingFlip_)
			flip(pixels_, w_, std::abs(h_), d_, type_);
		if (pendingSwapBR_)
			swapBR(pixels_, w_, std::abs(h_), d_, type_);
		pendingFlip_ = pendingSwapBR_ = false;
	}

	// 24-bit BITMAPINFOHEADER files only; rows are stored bottom-up in the file and come out
	// top-down; the B,G,R byte order of the file is NOT changed (codecs.cpp:306-317 never calls
	// swapBR), so a following writeJpg encodes blue as red -- the reference's behaviour.
	void Image::readBmp(std::string filepath, unsigned char** pixels, int& w, int& h, int& d, Type&)
	{
		std::ifstream f(filepath.c_str(), std::ios::binary);
		if (!f.is_open())
			throw std::runtime_error("Could not open .bmp file");

		unsigned char head[54];
		f.read(reinterpret_cast<char*>(head), sizeof(head));
		if (f.gcount() != (std::streamsize)sizeof(head) || head[0] != 'B' || head[1] != 'M')
			throw std::runtime_error("Could not parse .bmp file");

		auto le32 = [&](int off) {
			return (int32_t)((uint32_t)head[off] | ((uint32_t)head[off + 1] << 8) | ((uint32_t)head[off + 2] << 16) |
			                 ((uint32_t)head[off + 3] << 24));
		};
		const int biWidth = le32(18), biHeight = le32(22);
		const int rows = std::abs(biHeight);
		const int offset = biHeight > 0 ? 0 : rows - 1;
		const int padding = biWidth % 4;
		const size_t lenRow = (size_t)biWidth * 3;

		*pixels = new unsigned char[(size_t)rows * lenRow];
		for (int y = rows - 1; y >= 0; y--)
		{
			f.read(reinterpret_cast<char*>(*pixels + (size_t)std::abs(y - offset) * lenRow), (std::streamsize)lenRow);
			f.seekg(padding, std::ios::cur);
		}
		h = biHeight;
		w = biWidth;
		d = 3;
	}

	// codecs.cpp:821-849 with the GPU decoder where the reference calls njDecode / njGetImage; like the
	// reference, d is 3 whatever the file holds, so only colour files are read correctly
	void Image::readJpg(std::string filepath, unsigned char** pixels, int& w, int& h, int& d, Type&)
	{
		FILE* f = fopen(filepath.c_str(), "rb");
		if (!f)
			throw std::runtime_error("Error decoding the input file.\n");
		fseek(f, 0, SEEK_END);
		long size = ftell(f);
		fseek(f, 0, SEEK_SET);
		std::vector<unsigned char> file((size_t)std::max(size, 0L));
		size = (long)fread(file.data(), 1, file.size(), f);
		fclose(f);

		int nc = 0;
		if (!jpeg_gpu_decode_info(file.data(), (size_t)size, &w, &h, &nc))
			throw std::runtime_error("Error decoding the input file.\n");
		d = 3;
		*pixels = new unsigned char[(size_t)w * h * 3];
		if (!jpeg_gpu_decode(file.data(), (size_t)size, *pixels, (size_t)w * h * 3, &w, &h, &nc))
		{
			delete[] *pixels;
			*pixels = nullptr;
			throw std::runtime_error("Error decoding the input file.\n");
		}
	}

	// 24-bit BITMAPINFOHEADER, rows bottom-up, the reference's own row padding of (w % 4) bytes and
	// its bfSize formula (codecs.cpp:324-375) -- files round-trip through readBmp above
	void Image::writeBmp(std::string filepath, unsigned char* pixels, int& w, int& h, int& d, Type&)
	{
		std::ofstream f(filepath.c_str(), std::ios::binary);
		if (!f.is_open())
			throw std::runtime_error("Could not open .bmp file to write");

		unsigned char head[54] = {0};
		auto le32 = [&](int off, uint32_t v) {
			for (int b = 0; b < 4; ++b)
				head[off + b] = (unsigned char)(v >> (8 * b));
		};
		head[0] = 'B';
		head[1] = 'M';
		le32(2, (uint32_t)((3 * w + (w % 4)) * h));   // bfSize as the reference computes it
		le32(10, 54);                                  // bfOffBits
		le32(14, 40);                                  // biSize
		le32(18, (uint32_t)w);
		le32(22, (uint32_t)h);
		head[26] = 1;                                  // biPlanes
		head[28] = 24;                                 // biBitCount
		f.write(reinterpret_cast<const char*>(head), sizeof(head));

		const int rows = std::abs(h);
		const int offset = h > 0 ? 0 : rows - 1;
		const int padding = w % 4;
		const size_t lenRow = (size_t)w * d;
		for (int y = rows - 1; y >= 0; y--)
		{
			f.write(reinterpret_cast<const char*>(pixels + (size_t)std::abs(y - offset) * lenRow), (std::streamsize)lenRow);
			f.write("\0\0\0", padding);
		}
	}

	void Image::writeJpg(std::string filepath, unsigned char* pixels, int& w, int& h, int& d, Type&)
	{
		if (!pendingFlip_ && !pendingSwapBR_)
		{
			// codecs.cpp:853: tje_encode_to_file(filepath.c_str(), w, h, d, pixels) -- return value dropped
			jpeg_gpu_encode_to_file(filepath.c_str(), w, h, d, pixels);
			return;
		}
		// pending flip() / swapBR(): the same call with the operations folded into the kernel's pixel
		// loads -- bottom-up rows are a negative stride from the last row, B<->R is a load flag.
		// Quality 3 and the "0-byte file on failure" behaviour as in tje_encode_to_file (jpeg_enc.h:1177-1213).
		FILE* fd = fopen(filepath.c_str(), "wb");
		if (!fd)
			return;
		jpeg_gpu_image im = {};
		im.width = w;
		im.height = h;
		im.ncomp = d;
		im.quality_mode = JPEG_GPU_QMODE_TJE;
		im.quality = 3;
		im.subsampling = JPEG_GPU_SUB_444;
		im.flags = pendingSwapBR_ ? JPEG_GPU_FLAG_SWAP_RB : 0;
		const long long lenRow = (long long)w * d;
		im.stride = pendingFlip_ ? (int)-lenRow : (int)lenRow;
		im.pixels = pendingFlip_ && h > 0 ? pixels + (size_t)(h - 1) * (size_t)lenRow : pixels;
		std::vector<unsigned char> out(jpeg_gpu_max_encoded_size(w, h, d, JPEG_GPU_SUB_444));
		jpeg_gpu_output o = {out.data(), out.size(), 0, 0};
		if (!out.empty() && jpeg_gpu_encode_batch(&im, 1, &o, nullptr) == 1)
			fwrite(out.data(), 1, o.size, fd);
		fclose(fd);
	}
}
