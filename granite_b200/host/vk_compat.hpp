// vk_compat.hpp -- plain constants for the handful of Vk* names the render-graph declaration
// surface mentions (AttachmentInfo::format etc.).  Values equal Vulkan's so existing builder
// code compiles unchanged; nothing here depends on Vulkan headers.
#pragma once

#include <cstdint>

using VkFormat = int32_t;
constexpr VkFormat VK_FORMAT_UNDEFINED = 0;
constexpr VkFormat VK_FORMAT_R8_UNORM = 9;
constexpr VkFormat VK_FORMAT_R8G8_UNORM = 16;
constexpr VkFormat VK_FORMAT_R8G8B8A8_UNORM = 37;
constexpr VkFormat VK_FORMAT_R8G8B8A8_SRGB = 43;
constexpr VkFormat VK_FORMAT_A2B10G10R10_UNORM_PACK32 = 64;
constexpr VkFormat VK_FORMAT_R16G16_SFLOAT = 83;
constexpr VkFormat VK_FORMAT_R16G16B16A16_SFLOAT = 97;
constexpr VkFormat VK_FORMAT_B10G11R11_UFLOAT_PACK32 = 122;
constexpr VkFormat VK_FORMAT_D32_SFLOAT = 126;

using VkDeviceSize = uint64_t;
using VkFlags = uint32_t;
using VkBufferUsageFlags = VkFlags;
using VkImageUsageFlags = VkFlags;
using VkPipelineStageFlags2 = uint64_t;
using VkAccessFlags2 = uint64_t;
constexpr VkBufferUsageFlags VK_BUFFER_USAGE_TRANSFER_DST_BIT = 0x2;
constexpr VkBufferUsageFlags VK_BUFFER_USAGE_UNIFORM_BUFFER_BIT = 0x10;
constexpr VkBufferUsageFlags VK_BUFFER_USAGE_STORAGE_BUFFER_BIT = 0x20;
constexpr VkImageUsageFlags VK_IMAGE_USAGE_SAMPLED_BIT = 0x4;
// Stage / access masks the declaration surface takes (RenderPass::add_proxy_*, add_external_lock).  On this executor
// a pass is a unit of stream order, so the values only have to be non-zero where the reference asserts that.
constexpr VkPipelineStageFlags2 VK_PIPELINE_STAGE_2_PRE_RASTERIZATION_SHADERS_BIT = 0x4000000000ull;
constexpr VkPipelineStageFlags2 VK_PIPELINE_STAGE_FRAGMENT_SHADER_BIT = 0x80ull;
constexpr VkPipelineStageFlags2 VK_PIPELINE_STAGE_2_COMPUTE_SHADER_BIT = 0x800ull;
constexpr VkAccessFlags2 VK_ACCESS_2_SHADER_SAMPLED_READ_BIT = 0x100000000ull;
constexpr VkAccessFlags2 VK_ACCESS_2_SHADER_STORAGE_READ_BIT = 0x200000000ull;
constexpr VkAccessFlags2 VK_ACCESS_2_SHADER_STORAGE_WRITE_BIT = 0x400000000ull;

struct VkClearDepthStencilValue
{
	float depth;
	uint32_t stencil;
};

union VkClearColorValue
{
	float float32[4];
	int32_t int32[4];
	uint32_t uint32[4];
};

// VK_EXT_hdr_metadata, the fields setup_hdr10_pq_encoding reads (renderer/post/hdr.cpp:563-593, 652).
struct VkXYColorEXT
{
	float x, y;
};

struct VkHdrMetadataEXT
{
	VkXYColorEXT displayPrimaryRed, displayPrimaryGreen, displayPrimaryBlue, whitePoint;
	float maxLuminance, minLuminance, maxContentLightLevel, maxFrameAverageLightLevel;
};

namespace Granite
{
inline unsigned format_texel_size(VkFormat format)
{
	switch (format)
	{
	case VK_FORMAT_R8_UNORM:
		return 1;
	case VK_FORMAT_R8G8_UNORM:
		return 2;
	case VK_FORMAT_R16G16B16A16_SFLOAT:
		return 8;
	case VK_FORMAT_R8G8B8A8_UNORM:
	case VK_FORMAT_R8G8B8A8_SRGB:
	case VK_FORMAT_A2B10G10R10_UNORM_PACK32:
	case VK_FORMAT_R16G16_SFLOAT:
	case VK_FORMAT_B10G11R11_UFLOAT_PACK32:
	case VK_FORMAT_D32_SFLOAT:
		return 4;
	default:
		return 0;
	}
}

inline bool format_is_srgb(VkFormat format) { return format == VK_FORMAT_R8G8B8A8_SRGB; }
} // namespace Granite
