// hdr.hpp -- HDR post chain pass builders, same entry points as renderer/post/hdr.hpp:29-49.
#pragma once

#include <string>

#include "../render_context.hpp"
#include "../render_graph.hpp"

namespace Granite
{
class HDRDynamicExposureInterface
{
public:
	virtual ~HDRDynamicExposureInterface() = default;
	virtual float get_exposure() const = 0;
};

struct HDROptions
{
	bool dynamic_exposure = true;
};

// Bloom threshold -> 4x downsample (last with temporal feedback) -> average luminance ->
// 3x upsample as ONE "bloom-compute" pass, then a "tonemap" pass (renderer/post/hdr.cpp:308-400).
void setup_hdr_postprocess_compute(RenderGraph &graph, const FrameParameters &frame, const std::string &input, const std::string &output,
                                   const HDROptions &options, const HDRDynamicExposureInterface *iface = nullptr);

// The reference's fragment-shader variant (hdr.cpp:402-561) computes the same images with the
// same arithmetic; on this executor it is the same kernels, so it forwards.
void setup_hdr_postprocess(RenderGraph &graph, const FrameParameters &frame, const std::string &input, const std::string &output,
                           const HDROptions &options, const HDRDynamicExposureInterface *iface = nullptr);
} // namespace Granite
