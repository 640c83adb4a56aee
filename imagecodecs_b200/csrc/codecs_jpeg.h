// codecs_jpeg.h -- the JPEG-write slice of the reference's façade, host side, C++.
//
// Mirrors ImageCodecs::Image (codecs.h:16-103) for exactly the members the JPEG write path
// touches, with the same names, argument meaning and error behaviour, so that the reference's
// own usage (tests.cpp:98-108: read a fixture, write it back) reads the same:
//
//     ImageCodecs::Image img;
//     img.read("cat.bmp");          // codecs.cpp:255-320 semantics (B,G,R bytes kept as they are)
//     img.write("cat.jpg");         // codecs.cpp:106-107 -> writeJpg (:851-854)
//
// writeJpg calls jpeg_gpu_encode_to_file exactly where codecs.cpp:853 calls
// tje_encode_to_file; every other format of the reference is out of scope.
//
// flip() and swapBR() (codecs.h:78,99; codecs.cpp:162-251) do not touch the
// pixels here.  They are remembered, and a following write("x.jpg") hands them to the encoder as
// load-time swizzles (a negative row stride, JPEG_GPU_FLAG_SWAP_RB): the kernel reads the rows /
// channels in the other order, no host pass over the image happens.  Anything that LOOKS at the
// pixels (data(), write("x.bmp")) first brings the buffer into the state the reference would have
// left it in, so the difference is not observable.
#pragma once
#include <stdexcept>
#include <string>

namespace ImageCodecs
{
	enum class Type
	{
		UBYTE,
		USHORT,
		FLOAT
	};

	class __attribute__((visibility("default"))) Image
	{
		int h_ = 0;
		int w_ = 0;
		int d_ = 0;
		unsigned char* pixels_ = nullptr;
		Type type_ = Type::UBYTE;

		bool pendingFlip_ = false;    // flip() / swapBR() called an odd number of times since the pixels were last touched
		bool pendingSwapBR_ = false;

		void flip(unsigned char* pixels, const int w, const int h, const int d, const Type& type);
		void swapBR(unsigned char* pixels, const int w, const int h, const int d, const Type& type);
		void materialize();           // apply the pending operations to pixels_ (what the reference does at once)

		// codecs per filetype (the JPEG write path and its BMP feeder):
		void readBmp(std::string filename, unsigned char** pixels, int& w, int& h, int& d, Type& type);
		void readJpg(std::string filename, unsigned char** pixels, int& w, int& h, int& d, Type& type);
		void writeBmp(std::string filename, unsigned char* pixels, int& w, int& h, int& d, Type& type);
		void writeJpg(std::string filename, unsigned char* pixels, int& w, int& h, int& d, Type& type);

	public:
		inline int byteSize() { return type_ == Type::FLOAT ? 4 : (type_ == Type::USHORT ? 2 : 1); }
		inline int channels() { return d_; }
		inline int cols() { return w_; }
		inline unsigned char** data() { materialize(); return &pixels_; }
		inline bool empty() { return h_ == 0 || w_ == 0 || d_ == 0 || pixels_ == nullptr; }
		inline void flip() { pendingFlip_ = !pendingFlip_; }
		// adopts the pointer (the destructor delete[]s it), like codecs.h:89-95
		inline void load(unsigned char* pixels, int w, int h, int channels)
		{
			d_ = channels;
			w_ = w;
			h_ = h;
			pixels_ = pixels;
			pendingFlip_ = pendingSwapBR_ = false;
		}
		void read(std::string filepath);
		inline int rows() { return h_; }
		inline void swapBR() { pendingSwapBR_ = !pendingSwapBR_; }
		inline int totalBytes() { return w_ * h_ * d_ * byteSize(); }
		inline Type type() { return type_; }
		void write(std::string filepath);
		~Image() { delete[] pixels_; }
	};
}
