// RGBValue / Image / PPM writer -- same surface and arithmetic as the reference's framebuffer classes
// (main.cpp:21-73 RGBValue, :79-100 Image, :102-128 writeImage), written from scratch.
//   * RGBValue clamps each channel to [0,1]; NaN passes through both comparisons      (main.cpp:29-41)
//   * Image stores float RGB at 3*(W*j+i)+c                                           (main.cpp:88-94)
//   * writeImage: "P6\n%i %i\n255\n" then (unsigned char)(v*255.0f) per channel,
//     i.e. truncation toward zero, 1.0 -> 255; rows top to bottom; one fwrite         (main.cpp:112-119)
#pragma once
#include <cstdio>
#include <vector>

class RGBValue {
public:
    float r, b, g;
    RGBValue(float rI = 0, float gI = 0, float bI = 0) : r(rI), b(bI), g(gI) {
        if (r > 1) r = 1.0f;
        if (g > 1) g = 1.0f;
        if (b > 1) b = 1.0f;
        if (r < 0) r = 0.0f;
        if (g < 0) g = 0.0f;
        if (b < 0) b = 0.0f;
    }
    float operator[](int i) const { return i == 1 ? g : (i == 2 ? b : r); }
    float& operator[](int i) { return i == 1 ? g : (i == 2 ? b : r); }
};

class Image {
public:
    std::vector<float> _image;
    int _width, _height;
    Image(int width, int height) : _image((size_t)3 * width * height), _width(width), _height(height) {}
    void setPixel(int i, int j, const RGBValue& rgb) {
        float* px = &_image[3 * ((size_t)_width * j + i)];
        px[0] = rgb[0]; px[1] = rgb[1]; px[2] = rgb[2];
    }
    bool writeImage(const char* filename) {
        FILE* f = fopen(filename, "wb");
        if (!f) { printf("dump file problem... file\n"); return false; }
        fprintf(f, "P6\n%i %i\n255\n", _width, _height);
        std::vector<unsigned char> bytes(_image.size());
        for (size_t i = 0; i < _image.size(); ++i) bytes[i] = (unsigned char)(_image[i] * 255.0f);
        size_t ok = fwrite(bytes.data(), (size_t)_width * _height * 3, 1, f);
        fclose(f);
        if (ok != 1) { printf("Dump file problem... fwrite\n"); return false; }
        return true;
    }
};
