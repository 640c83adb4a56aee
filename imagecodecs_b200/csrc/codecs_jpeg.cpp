// codecs_jpeg.cpp -- see codecs_jpeg.h.  Host C++ in the reference's style; the only call
// into the GPU library is the drop-in twin of tje_encode_to_file.
#include "codecs_jpeg.h"

#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <fstream>

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
		if (ext == ".bmp")
			readBmp(filepath, &pixels_, w_, h_, d_, type_);
		else
			throw std::invalid_argument("Cannot parse filetype");

		if (pixels_ == nullptr)
			throw std::runtime_error("Could not read image data");
	}

	void Image::write(std::string filepath)
	{
		// codecs.cpp:91-122: lower-cased extension picks the codec
		const std::string ext = lowerExtension(filepath);
		if (ext == ".jpg" || ext == ".jpeg")
			writeJpg(filepath, pixels_, w_, h_, d_, type_);
		else
			throw std::invalid_argument("Cannot parse filetype");
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

	void Image::writeJpg(std::string filepath, unsigned char* pixels, int& w, int& h, int& d, Type&)
	{
		// codecs.cpp:853: tje_encode_to_file(filepath.c_str(), w, h, d, pixels) -- return value dropped
		jpeg_gpu_encode_to_file(filepath.c_str(), w, h, d, pixels);
	}
}
