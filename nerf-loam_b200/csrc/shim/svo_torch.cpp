// TorchScript custom-class shim: registers `torch.classes.svo.Octree` (and `torch.ops.svo.encode`) with the interface of
// the reference's extension (third_party/sparse_octree/src/bindings.cpp:4-33, include/octree.h:65-122), implemented on the
// C ABI of libnerfloam_b200 (nl_octree_*, include/nerfloam_b200.h section 1).  With this library the one line of the
// reference that loads its octree,
//     torch.classes.load_library(".../svo.cpython-38-x86_64-linux-gnu.so")        (src/mapping.py:19-20)
// only needs the new path; `torch.classes.svo.Octree()` and everything called on it stay as they are.
// Host-only code (the octree lives on the CPU in the reference too); built by nerf-loam_b200/build.py with g++.
#include <torch/custom_class.h>
#include <torch/script.h>

#include <stdexcept>
#include <tuple>
#include <vector>

#include "nerfloam_b200.h"

namespace {

struct Octree : torch::CustomClassHolder {
    Octree() = default;
    // bindings.cpp:27-30 (__setstate__): re-create and replay every inserted tensor
    Octree(int64_t grid_dim, int64_t feat_dim, double voxel_size, std::vector<torch::Tensor> all_pts) {
        init(grid_dim, feat_dim, voxel_size);
        for (auto &p : all_pts) insert(p);
    }
    ~Octree() override {
        if (h_) nl_octree_destroy(h_);
    }
    Octree(const Octree &) = delete;
    Octree &operator=(const Octree &) = delete;

    void init(int64_t grid_dim, int64_t feat_dim, double voxel_size) {   // octree.cpp:34-49
        if (h_) nl_octree_destroy(h_);
        h_ = nl_octree_create(grid_dim, feat_dim, voxel_size);
        if (!h_) throw std::runtime_error(nl_last_error());
        size_ = grid_dim; feat_dim_ = feat_dim; voxel_size_ = voxel_size;
        all_pts.clear();
    }
    // int32 [N,3] voxel coordinates on the CPU (the reference reads them through accessor<int,2>)
    static torch::Tensor as_i32(const torch::Tensor &t) {
        TORCH_CHECK(t.scalar_type() == torch::kInt32, "expected an int32 tensor");
        return t.detach().cpu().contiguous();
    }
    void insert(torch::Tensor vox) {                                     // octree.cpp:51-111
        if (!h_) { printf("Octree not initialized!\n"); return; }       // octree.cpp:56-59
        if (vox.dim() != 2 || vox.size(1) != 3) {                        // octree.cpp:62-66
            printf("Point dimensions mismatch: inputs are %ld expect 3\n", (long)(vox.dim() ? vox.size(-1) : 0));
            return;
        }
        auto v = as_i32(vox);
        if (nl_octree_insert(h_, v.data_ptr<int32_t>(), v.size(0)) != NL_OK) throw std::runtime_error(nl_last_error());
        all_pts.push_back(v);
    }
    double try_insert(torch::Tensor pts) {                               // octree.cpp:113-149
        if (!h_ || pts.dim() != 2 || pts.size(1) != 3) return -1.0;
        auto v = as_i32(pts);
        return nl_octree_try_insert(h_, v.data_ptr<int32_t>(), v.size(0));
    }
    bool has_voxel(torch::Tensor pose) {                                 // octree.cpp:173-206
        if (!h_ || pose.numel() != 3) return false;
        auto v = as_i32(pose.reshape({-1}));
        return nl_octree_has_voxel(h_, v.data_ptr<int32_t>()) != 0;
    }
    torch::Tensor get_features(torch::Tensor) { return torch::Tensor(); }   // empty body in the reference (octree.cpp:208-210)
    torch::Tensor get_voxels() {                                         // octree.cpp:228-252
        check();
        const int64_t n = nl_octree_count_nodes(h_);
        auto out = torch::empty({n, 4}, torch::kFloat32);
        const int64_t rows = nl_octree_get_voxels(h_, out.data_ptr<float>(), n);
        return out.narrow(0, 0, rows);
    }
    torch::Tensor get_leaf_voxels() {                                    // octree.cpp:212-226
        check();
        const int64_t n = nl_octree_count_leaf_nodes(h_);
        auto out = torch::empty({n > 0 ? n : 1, 3}, torch::kFloat32);
        const int64_t rows = nl_octree_get_leaf_voxels(h_, out.data_ptr<float>(), n);
        return out.narrow(0, 0, rows);
    }
    int64_t count_nodes() { check(); return nl_octree_count_nodes(h_); }             // octree.cpp:344-364
    int64_t count_leaf_nodes() { check(); return nl_octree_count_leaf_nodes(h_); }   // octree.cpp:366-389
    std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> get_centres_and_children() {   // octree.cpp:293-342
        check();
        const int64_t n = nl_octree_count_export_nodes(h_);
        auto voxels = torch::empty({n, 4}, torch::kFloat32), children = torch::empty({n, 8}, torch::kFloat32);
        auto features = torch::empty({n, 8}, torch::kInt32);
        if (nl_octree_export(h_, voxels.data_ptr<float>(), children.data_ptr<float>(), features.data_ptr<int32_t>()) != NL_OK)
            throw std::runtime_error(nl_last_error());
        return std::make_tuple(voxels, children, features);
    }
    void check() const {
        if (!h_) throw std::runtime_error("Octree not initialized!");
    }

    nl_octree *h_ = nullptr;
    int64_t size_ = 0, feat_dim_ = 0;
    double voxel_size_ = 0.0;
    std::vector<torch::Tensor> all_pts;
};

int64_t encode_torch(torch::Tensor pt) {                                 // utils.h:106-109 via test.h
    auto v = pt.detach().cpu().to(torch::kInt32).contiguous().reshape({-1});
    TORCH_CHECK(v.numel() == 3, "encode expects 3 coordinates");
    const int32_t *p = v.data_ptr<int32_t>();
    return (int64_t)nl_morton_encode(p[0], p[1], p[2]);
}

}  // namespace

// Serialised form used by torch.save / copy.deepcopy of the reference's class: tree extent, feature width, voxel size and
// every tensor that was ever inserted (the tree is rebuilt by replaying them, which also reproduces the node ids).
using OctreeState = std::tuple<int64_t, int64_t, double, std::vector<torch::Tensor>>;

OctreeState octree_getstate(const c10::intrusive_ptr<Octree> &tree) {
    return OctreeState(tree->size_, tree->feat_dim_, tree->voxel_size_, tree->all_pts);
}

c10::intrusive_ptr<Octree> octree_setstate(OctreeState st) {
    auto &[extent, width, voxel, inserted] = st;
    return c10::make_intrusive<Octree>(extent, width, voxel, std::move(inserted));
}

TORCH_LIBRARY(svo, lib) {
    lib.def("encode", &encode_torch);
    auto cls = lib.class_<Octree>("Octree");
    cls.def(torch::init<>());
    // allocation and queries, names and argument lists as in third_party/sparse_octree/src/bindings.cpp:11-22
    cls.def("init", &Octree::init).def("insert", &Octree::insert).def("try_insert", &Octree::try_insert);
    cls.def("has_voxel", &Octree::has_voxel).def("get_features", &Octree::get_features);
    cls.def("get_voxels", &Octree::get_voxels).def("get_leaf_voxels", &Octree::get_leaf_voxels);
    cls.def("count_nodes", &Octree::count_nodes).def("count_leaf_nodes", &Octree::count_leaf_nodes);
    cls.def("get_centres_and_children", &Octree::get_centres_and_children);
    // def_pickle only takes lambdas (torch/custom_class.h)
    cls.def_pickle([](const c10::intrusive_ptr<Octree> &tree) { return octree_getstate(tree); },
                   [](OctreeState st) { return octree_setstate(std::move(st)); });
}
