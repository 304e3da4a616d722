// fxaa.hpp -- renderer/post/fxaa.hpp
#pragma once

#include <string>

#include "../render_graph.hpp"

namespace Granite
{
void setup_fxaa_postprocess(RenderGraph &graph, const std::string &input, const std::string &output, VkFormat output_format = VK_FORMAT_UNDEFINED);
}
