// tests/cpp/write_jpg_like_reference.cpp -- what tests.cpp:98-108 does for one file, through
// the C++ host layer:   ./a.out in.bmp out.jpg        (read a fixture, write it as JPEG)
//                or:    ./a.out --raw w h d in.raw out.jpg
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
