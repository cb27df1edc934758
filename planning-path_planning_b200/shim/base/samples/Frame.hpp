// Stand-in for base/samples/Frame.hpp: an 8-bit image container exposing the
// five members the DyMu local layer reads (getHeight/getWidth/getRowSize/
// getPixelSize/image; /root/reference/src/DyMu_LocalPathRepairing.cpp:206-244).
#ifndef DYMU_SHIM_BASE_SAMPLES_FRAME_HPP
#define DYMU_SHIM_BASE_SAMPLES_FRAME_HPP
#include <cstdint>
#include <vector>
#include <base/Time.hpp>
namespace base
{
namespace samples
{
namespace frame
{
struct Frame
{
    std::vector<uint8_t> image;
    uint16_t width, height;
    uint32_t row_size, pixel_size;
    Frame() : width(0), height(0), row_size(0), pixel_size(1) {}
    Frame(uint16_t w, uint16_t h, uint32_t pixel_bytes = 1)
        : image(static_cast<std::size_t>(w) * h * pixel_bytes, 0),
          width(w),
          height(h),
          row_size(w * pixel_bytes),
          pixel_size(pixel_bytes)
    {
    }
    uint16_t getHeight() const { return height; }
    uint16_t getWidth() const { return width; }
    uint32_t getRowSize() const { return row_size; }
    uint32_t getPixelSize() const { return pixel_size; }
};
}  // namespace frame
}  // namespace samples
}  // namespace base
#endif
