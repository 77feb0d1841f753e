// Declaration-only stand-in for png++ (an external library that is absent from this image), so that
// the reference's OWN header external/deval_lib/src/evaluate_depth.h compiles unmodified from where
// it lies.  Only the reference's file-IO paths (readDepthMap / write* / errorImage / imageFormat /
// eval) touch these types; oracle/ref_deval/shim.cpp never calls them.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <cstdint>
#include <istream>
#include <string>
namespace png {
enum color_type { color_type_gray = 0, color_type_rgb = 2 };
struct rgb_pixel {
    unsigned char red, green, blue;
    rgb_pixel(unsigned char r = 0, unsigned char g = 0, unsigned char b = 0) : red(r), green(g), blue(b) {}
};
typedef uint16_t gray_pixel_16;
template <class P>
class image {
  public:
    image() : w_(0), h_(0) {}
    image(size_t w, size_t h) : w_(w), h_(h) {}
    explicit image(const std::string&) : w_(0), h_(0) {}
    size_t get_width() const { return w_; }
    size_t get_height() const { return h_; }
    P get_pixel(size_t, size_t) const { return P(); }
    void set_pixel(size_t, size_t, P) {}
    void write(const std::string&) {}
  private:
    size_t w_, h_;
};
template <class S>
class reader {
  public:
    explicit reader(S&) {}
    void read_info() {}
    color_type get_color_type() const { return color_type_gray; }
    size_t get_bit_depth() const { return 0; }
    size_t get_width() const { return 0; }
    size_t get_height() const { return 0; }
};
}  // namespace png
