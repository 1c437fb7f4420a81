// TEST INFRASTRUCTURE (oracle). Headless replacement for the reference's src/utils/gui.cpp, which
// needs CImg + X11 headers that this image does not have. Declarations: src/utils/gui.h:11-34.
#include "gui.h"

#include <string>

namespace cimg_library { class CImgDisplay {}; }

void gui::display(const std::uint8_t*, int, int, int) {}

dynamic_gui_impl::dynamic_gui_impl(int w, int h, int, const std::string&) : width(w), height(h) {}
dynamic_gui_impl::~dynamic_gui_impl() {}
template<typename T> void dynamic_gui_impl::show(const T*) {}
template void dynamic_gui_impl::show<std::uint8_t>(const std::uint8_t*);
template void dynamic_gui_impl::show<std::int32_t>(const std::int32_t*);
void dynamic_gui_impl::_display_thread() {}
