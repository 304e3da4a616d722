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

// HDR10 output encoding (renderer/post/hdr.hpp:51-59): a "pq10" pass that composites the scene colour with the UI
// layer, converts Rec.709 to the display's primaries and writes ST.2084 code values into an A2B10G10R10 image.
struct HDR10PQEncodingConfig
{
	float hdr_pre_exposure;
	float ui_pre_exposure;
};
void setup_hdr10_pq_encoding(RenderGraph &graph, const std::string &output, const std::string &hdr_input, const std::string &ui_input,
                             const HDR10PQEncodingConfig &config, const VkHdrMetadataEXT &static_metadata);
// Rec.709 -> display primaries (hdr.cpp:580-593): inverse(XYZ(display)) * XYZ(Rec.709), D65 assumed as Vulkan does.
muglm::mat4 compute_rec709_to_display_primaries(const VkHdrMetadataEXT &metadata);
} // namespace Granite
