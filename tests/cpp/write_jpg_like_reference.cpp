// tests/cpp/write_jpg_like_reference.cpp -- what tests.cpp:98-108 does for one file, through
// the C++ host layer:   ./a.out in.bmp out.jpg        (read a fixture, write it as JPEG)
//                or:    ./a.out --raw w h d in.raw out.jpg
//                or:    ./a.out --ops <f|s|fs|sf...> in.bmp out.jpg|out.bmp   (f = img.flip(), s = img.swapBR() before the write)
//                or:    ./a.out --ops-peek <ops> in.bmp out.jpg               (same, but data() is looked at before the write)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <iostream>

#include "codecs_jpeg.h"

int main(int argc, char** argv)
{
	try
	{
		ImageCodecs::Image img;
		if (argc == 7 && !strcmp(argv[1], "--raw"))
		{
			const int w = atoi(argv[2]), h = atoi(argv[3]), d = atoi(argv[4]);
			unsigned char* px = new unsigned char[(size_t)w * h * d];
			FILE* f = fopen(argv[5], "rb");
			if (!f || fread(px, 1, (size_t)w * h * d, f) != (size_t)w * h * d) return 2;
			fclose(f);
			img.load(px, w, h, d);
			img.write(argv[6]);
		}
		else if (argc == 5 && (!strcmp(argv[1], "--ops") || !strcmp(argv[1], "--ops-peek")))
		{
			img.read(argv[3]);
			for (const char* c = argv[2]; *c; ++c)
			{
				if (*c == 'f') img.flip();
				else if (*c == 's') img.swapBR();
			}
			if (!strcmp(argv[1], "--ops-peek") && *img.data() == nullptr) return 3;
			img.write(argv[4]);
		}
		else if (argc == 3)
		{
			img.read(argv[1]);
			img.write(argv[2]);
		}
		else
			return 64;
		std::cout << img.cols() << "x" << img.rows() << "x" << img.channels() << std::endl;
	}
	catch (std::exception& e)
	{
		std::cerr << e.what() << std::endl;
		return 1;
	}
	return EXIT_SUCCESS;
}
