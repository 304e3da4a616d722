#include "nccl_collectives.hpp"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

namespace Granite
{
namespace
{
// Minimal NCCL ABI (stable across 2.x): opaque comm, 128-byte unique id, enums as ints.
struct NcclUniqueId
{
	char internal[NcclUniqueIdBytes];
};
using ncclComm_t = void *;
constexpr int ncclSuccess = 0;
constexpr int ncclInt8 = 0;   // ncclChar
constexpr int ncclFloat32 = 7; // ncclFloat
constexpr int ncclSum = 0;

struct NcclApi
{
	void *handle = nullptr;
	int (*GetUniqueId)(NcclUniqueId *) = nullptr;
	int (*CommInitRank)(ncclComm_t *, int, NcclUniqueId, int) = nullptr;
	int (*CommDestroy)(ncclComm_t) = nullptr;
	int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, void *) = nullptr;
	int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, void *) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
	std::string error;
};

NcclApi &api()
{
	static NcclApi a;
	static std::once_flag once;
	std::call_once(once, [] {
		const char *names[] = { "libnccl.so.2", "libnccl.so" };
		for (auto *n : names)
		{
			a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
			if (a.handle)
				break;
		}
		if (!a.handle)
		{
			a.error = "libnccl.so.2 not found (import torch first, or add NCCL to LD_LIBRARY_PATH)";
			return;
		}
#define GRB_SYM(field, name)                                              \
	a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name)); \
	if (!a.field)                                                         \
		a.error = std::string("missing NCCL symbol ") + name;
		GRB_SYM(GetUniqueId, "ncclGetUniqueId")
		GRB_SYM(CommInitRank, "ncclCommInitRank")
		GRB_SYM(CommDestroy, "ncclCommDestroy")
		GRB_SYM(AllReduce, "ncclAllReduce")
		GRB_SYM(Broadcast, "ncclBroadcast")
		GRB_SYM(GroupStart, "ncclGroupStart")
		GRB_SYM(GroupEnd, "ncclGroupEnd")
		GRB_SYM(GetErrorString, "ncclGetErrorString")
#undef GRB_SYM
	});
	return a;
}

bool nccl_ok(int rc, const char *what)
{
	if (rc == ncclSuccess)
		return true;
	Vulkan::log_error("%s: %s\n", what, api().GetErrorString ? api().GetErrorString(rc) : "NCCL error");
	return false;
}
} // namespace

NcclCollectives::~NcclCollectives()
{
	if (comm && api().CommDestroy)
		api().CommDestroy(comm);
}

bool NcclCollectives::get_unique_id(unsigned char out[NcclUniqueIdBytes], std::string &error)
{
	auto &a = api();
	if (!a.error.empty())
	{
		error = a.error;
		return false;
	}
	NcclUniqueId id;
	if (a.GetUniqueId(&id) != ncclSuccess)
	{
		error = "ncclGetUniqueId failed";
		return false;
	}
	std::memcpy(out, id.internal, NcclUniqueIdBytes);
	return true;
}

bool NcclCollectives::init(const unsigned char id_bytes[NcclUniqueIdBytes], unsigned rank_, unsigned world_size, std::string &error)
{
	auto &a = api();
	if (!a.error.empty())
	{
		error = a.error;
		return false;
	}
	NcclUniqueId id;
	std::memcpy(id.internal, id_bytes, NcclUniqueIdBytes);
	int rc = a.CommInitRank(&comm, (int)world_size, id, (int)rank_);
	if (rc != ncclSuccess)
	{
		error = std::string("ncclCommInitRank: ") + a.GetErrorString(rc);
		comm = nullptr;
		return false;
	}
	rank = rank_;
	world = world_size;
	return true;
}

bool NcclCollectives::all_gather_rows(Vulkan::CommandBuffer &cmd, Vulkan::ImageView &image, const std::vector<GrbRows> &rows)
{
	if (!comm || rows.size() != world)
		return false;
	// Bands differ in height, so this is a grouped set of broadcasts (one root per band), which
	// NCCL fuses into a single launch over NVLink.
	auto &a = api();
	auto *base = static_cast<unsigned char *>(image.get_image().get_device_pointer());
	const size_t pitch = image.get_image().get_row_pitch();
	bool ok = nccl_ok(a.GroupStart(), "ncclGroupStart");
	for (unsigned r = 0; r < world && ok; r++)
	{
		size_t bytes = (size_t)(rows[r].y1 - rows[r].y0) * pitch;
		void *p = base + (size_t)rows[r].y0 * pitch;
		ok = nccl_ok(a.Broadcast(p, p, bytes, ncclInt8, (int)r, comm, cmd.get_stream_handle()), "ncclBroadcast");
	}
	ok = nccl_ok(a.GroupEnd(), "ncclGroupEnd") && ok;
	return ok;
}

bool NcclCollectives::all_reduce_sum(Vulkan::CommandBuffer &cmd, float *data, size_t count)
{
	if (!comm)
		return false;
	return nccl_ok(api().AllReduce(data, data, count, ncclFloat32, ncclSum, comm, cmd.get_stream_handle()), "ncclAllReduce");
}
} // namespace Granite
