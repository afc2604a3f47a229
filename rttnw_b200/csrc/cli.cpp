// cli.cpp — `rttnw <scene>`: the reference's CLI (src/main.rs:236-258) on top of the C ABI.
// Same positional argument, same usage text, same stdout lines, writes image.png (RGBA8) in the
// CWD and reads assets/earth.png relative to it. Optional flags default to the reference values:
//   --spp N  --width W  --height H  --seed S  --scene-seed S  --gpus N  --out PATH  --chunk N
//   --checkpoint PATH [--checkpoint-every K]  --resume PATH  --stop-after-chunks N
// Checkpoint / resume (the reference keeps pixels in RAM until the final save_buffer, main.rs:202-231, so a
// killed 10k-spp render loses everything): after every K chunks each GPU writes its fp32 accumulator and the
// next sample index it would render to PATH.<gpu>; --resume reloads them (same scene, size, spp, seed, gpus)
// and carries on. Samples are keyed by their global index, so a resumed image equals an uninterrupted one up
// to fp32 summation order.
// With --gpus N the samples are sharded by global sample index, one PROCESS per device, each seeing ONLY its device
// (CUDA_VISIBLE_DEVICES is narrowed before the process touches CUDA): on an 8-GPU box the CUDA runtime takes 6-7 s to
// initialise when it can see every device, whether one context is created or eight, in one process or in eight
// (profiles/r2_multi_gpu_8.txt) — that, not the render, was 7.5 s of round 1's 9 s. The parent forks ranks 1..N-1
// before it touches CUDA and renders rank 0's share itself. Processes that cannot see each other's devices cannot map
// each other's memory, so the end-of-frame combine goes through the host: every child copies its accumulator (10 MB at
// 800x800) into a shared mapping and exits; the parent uploads them and runs the same fused sum + tonemap kernel on local
// copies (a few milliseconds per frame). RTTNW_SINGLE_PROCESS=1 keeps one process with a thread per device and the
// peer-mapped combine over NVLink; the NCCL combine (rtx_comm_*, rtx_accum_reduce) is what bench.py uses across ranks.
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "rttnw_b200.h"

static void usage(const char* argv0) {  // main.rs:239-249
    std::fprintf(stderr, "Usage: %s <scene>\n", argv0);
    std::fprintf(stderr, "Possible scenes:\n");
    static const char* names[] = {"random_scene", "two_spheres", "two_perlin_spheres", "earth", "simple_light",
                                  "empty_cornell_box", "cornell_box", "smoke_cornell_box", "final_scene"};
    for (int i = 0; i < 9; ++i) std::fprintf(stderr, "\t- %d: %s\n", i + 1, names[i]);
}

#define RTX(call)                                                            \
    do {                                                                     \
        if ((call) != RTX_OK) {                                              \
            std::fprintf(stderr, "%s failed: %s\n", #call, rtx_last_error()); \
            return 1;                                                        \
        }                                                                    \
    } while (0)

struct Rank {
    rtx_ctx* ctx = nullptr;
    rtx_scene* scene = nullptr;
    float* accum = nullptr;
    int rc = 0;
};

struct CheckpointHeader {  // little-endian, followed by width * height float4
    char magic[8];         // "RTXACC1"
    int32_t scene, width, height, spp_total, gpus, rank, next_sample, max_depth;
    uint64_t seed, scene_seed;
};

static bool write_checkpoint(const std::string& path, const CheckpointHeader& h, const std::vector<float>& acc) {
    std::string tmp = path + ".tmp";
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) return false;
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 && std::fwrite(acc.data(), sizeof(float), acc.size(), f) == acc.size();
    ok = std::fclose(f) == 0 && ok;
    return ok && std::rename(tmp.c_str(), path.c_str()) == 0;  // never leaves a torn file behind
}
static bool read_checkpoint(const std::string& path, CheckpointHeader& h, std::vector<float>& acc) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    bool ok = std::fread(&h, sizeof(h), 1, f) == 1 && std::memcmp(h.magic, "RTXACC1", 8) == 0 && h.width > 0 && h.height > 0;
    if (ok) {
        acc.resize((size_t)h.width * h.height * 4);
        ok = std::fread(acc.data(), sizeof(float), acc.size(), f) == acc.size();
    }
    std::fclose(f);
    return ok;
}

// Narrows CUDA_VISIBLE_DEVICES to the `rank`-th device this process could see (the rank-th entry of the variable when the
// user set one, else device `rank`). Must run before the first CUDA call of the process; the device is then number 0.
static void see_only_device(int rank) {
    std::string pick = std::to_string(rank);
    if (const char* cur = std::getenv("CUDA_VISIBLE_DEVICES")) {
        std::string list = cur;
        size_t pos = 0;
        for (int i = 0; i < rank && pos != std::string::npos; ++i) {
            pos = list.find(',', pos);
            if (pos != std::string::npos) ++pos;
        }
        if (pos == std::string::npos || pos >= list.size()) return;  // fewer entries than ranks: leave it, rtx_ctx_create will say so
        size_t end = list.find(',', pos);
        pick = list.substr(pos, end == std::string::npos ? std::string::npos : end - pos);
    }
    setenv("CUDA_VISIBLE_DEVICES", pick.c_str(), 1);
}

int main(int argc, char** argv) {
    int scene = -1, spp = -1, width = -1, height = -1, gpus = 1, chunk = 256, ckpt_every = 1, stop_after = -1;
    std::string ckpt_path, resume_path;
    unsigned long long seed = 1, scene_seed = 0;
    bool have_scene_seed = false;
    std::string out = "image.png";
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", name); std::exit(1); }
            return argv[++i];
        };
        if (a == "--spp") spp = std::atoi(next("--spp"));
        else if (a == "--width") width = std::atoi(next("--width"));
        else if (a == "--height") height = std::atoi(next("--height"));
        else if (a == "--gpus") gpus = std::atoi(next("--gpus"));
        else if (a == "--chunk") chunk = std::atoi(next("--chunk"));
        else if (a == "--seed") seed = std::strtoull(next("--seed"), nullptr, 0);
        else if (a == "--scene-seed") { scene_seed = std::strtoull(next("--scene-seed"), nullptr, 0); have_scene_seed = true; }
        else if (a == "--out") out = next("--out");
        else if (a == "--checkpoint") ckpt_path = next("--checkpoint");
        else if (a == "--checkpoint-every") ckpt_every = std::atoi(next("--checkpoint-every"));
        else if (a == "--resume") resume_path = next("--resume");
        else if (a == "--stop-after-chunks") stop_after = std::atoi(next("--stop-after-chunks"));
        else if (scene < 0 && !a.empty() && a[0] != '-') {
            char* end = nullptr;
            long v = std::strtol(a.c_str(), &end, 10);
            if (*end != 0) { std::fprintf(stderr, "Error: There was an error\n"); return 1; }  // parse() failure -> DummyError
            scene = (int)v;
        } else { usage(argv[0]); std::fprintf(stderr, "Error: There was an error\n"); return 1; }
    }
    if (scene < 0) { usage(argv[0]); std::fprintf(stderr, "Error: There was an error\n"); return 1; }
    std::printf("Scene number: %d\n", scene);
    auto t0 = std::chrono::steady_clock::now();
    rtx_scene_defaults def;
    if (rtx_builtin_scene_defaults(scene, &def) != RTX_OK) {
        std::fprintf(stderr, "%s\nError: There was an error\n", rtx_last_error());  // main.rs:179-182
        return 1;
    }
    std::printf("Running scene %s\n", def.name);
    if (spp < 0) spp = def.samples;
    if (width < 0) width = def.width;
    if (height < 0) height = def.height;
    if (gpus < 1) gpus = 1;
    if (chunk < 1) chunk = 1;
    if (!have_scene_seed) scene_seed = 0x5254544E57ull + (unsigned long long)scene;
    rtx_scene_desc* desc = nullptr;
    RTX(rtx_builtin_scene(scene, scene_seed, nullptr, &desc));

    const bool verbose = std::getenv("RTTNW_VERBOSE") != nullptr;
    auto lap = [&](const char* what) {
        if (verbose) std::fprintf(stderr, "[%8.3f s] %s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), what);
    };
    lap("scene description built");
    const bool one_device_per_process = std::getenv("RTTNW_SINGLE_PROCESS") == nullptr;
    std::vector<Rank> ranks((size_t)gpus);
    auto worker = [&](int r) {
        Rank& k = ranks[(size_t)r];
        auto chk = [&](int rc) { if (rc != RTX_OK && k.rc == 0) { k.rc = rc; std::fprintf(stderr, "gpu %d: %s\n", r, rtx_last_error()); } return rc == RTX_OK; };
        if (!chk(rtx_ctx_create(one_device_per_process ? 0 : r, nullptr, &k.ctx))) return;
        if (r == 0) lap("context created");
        if (!chk(rtx_scene_create(k.ctx, desc, &k.scene))) return;
        if (r == 0) lap("scene uploaded");
        size_t bytes = (size_t)width * height * 4 * sizeof(float);
        if (!chk(rtx_malloc(k.ctx, bytes, (void**)&k.accum))) return;
        if (!chk(rtx_memset_zero(k.ctx, k.accum, bytes))) return;
        int begin = (int)((long long)r * spp / gpus), end = (int)((long long)(r + 1) * spp / gpus);
        CheckpointHeader hdr;
        std::memset(&hdr, 0, sizeof(hdr));
        std::memcpy(hdr.magic, "RTXACC1", 8);
        hdr.scene = scene; hdr.width = width; hdr.height = height; hdr.spp_total = spp; hdr.gpus = gpus; hdr.rank = r;
        hdr.max_depth = def.max_depth; hdr.seed = seed; hdr.scene_seed = scene_seed;
        std::vector<float> host_acc;
        if (!resume_path.empty()) {
            CheckpointHeader got;
            std::string path = resume_path + "." + std::to_string(r);
            if (!read_checkpoint(path, got, host_acc) || got.scene != scene || got.width != width || got.height != height ||
                got.spp_total != spp || got.gpus != gpus || got.rank != r || got.seed != seed || got.scene_seed != scene_seed ||
                got.max_depth != def.max_depth || got.next_sample < begin || got.next_sample > end) {
                std::fprintf(stderr, "gpu %d: %s is not a checkpoint of this render\n", r, path.c_str());
                k.rc = 1;
                return;
            }
            if (!chk(rtx_memcpy_h2d(k.ctx, k.accum, host_acc.data(), bytes))) return;
            begin = got.next_sample;
        }
        int chunks_done = 0;
        for (int b = begin; b < end; b += chunk) {
            rtx_render_params p;
            std::memset(&p, 0, sizeof(p));
            p.width = width; p.height = height; p.spp_begin = b; p.spp_count = (end - b < chunk) ? end - b : chunk;
            p.max_depth = def.max_depth; p.seed = seed;
            if (!chk(rtx_render(k.ctx, k.scene, &p, k.accum, nullptr))) return;
            ++chunks_done;
            if (!ckpt_path.empty() && (chunks_done % (ckpt_every < 1 ? 1 : ckpt_every) == 0 || b + chunk >= end)) {
                host_acc.resize(bytes / sizeof(float));
                if (!chk(rtx_memcpy_d2h(k.ctx, host_acc.data(), k.accum, bytes))) return;  // ordered after the render on the ctx stream
                hdr.next_sample = b + p.spp_count;
                if (!write_checkpoint(ckpt_path + "." + std::to_string(r), hdr, host_acc)) {
                    std::fprintf(stderr, "gpu %d: cannot write checkpoint %s\n", r, ckpt_path.c_str());
                    k.rc = 1;
                    return;
                }
            }
            if (stop_after >= 0 && chunks_done >= stop_after) { k.rc = 3; return; }  // simulated interruption (tests)
        }
        chk(rtx_ctx_sync(k.ctx));
        if (r == 0) lap("render finished");
    };
    const bool multi_process = gpus > 1 && one_device_per_process;
    std::vector<const float*> peers;
    if (!multi_process && one_device_per_process) see_only_device(0);
    if (multi_process) {
        const size_t frame_bytes = (size_t)width * height * 4 * sizeof(float);
        // one slot per child for its accumulator, shared with the parent; a pipe per child for its status
        float* shared = (float*)mmap(nullptr, frame_bytes * (size_t)(gpus - 1), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
        if (shared == MAP_FAILED) { std::perror("mmap"); return 1; }
        struct Child {
            pid_t pid = -1;
            int up = -1;
        };
        std::vector<Child> children;
        std::fflush(stdout);
        std::fflush(stderr);
        for (int r = 1; r < gpus; ++r) {
            int up[2];
            if (pipe(up) != 0) { std::perror("pipe"); return 1; }
            pid_t pid = fork();
            if (pid < 0) { std::perror("fork"); return 1; }
            if (pid == 0) {  // rank r: sees its device only, renders its share, leaves the accumulator in the shared mapping
                close(up[0]);
                see_only_device(r);
                worker(r);
                Rank& k = ranks[(size_t)r];
                int32_t rc = k.rc;
                if (rc == 0 && rtx_memcpy_d2h(k.ctx, (char*)shared + frame_bytes * (size_t)(r - 1), k.accum, frame_bytes) != RTX_OK) {
                    std::fprintf(stderr, "gpu %d: %s\n", r, rtx_last_error());
                    rc = 1;
                }
                ssize_t w = write(up[1], &rc, sizeof(rc));
                (void)w;
                _exit(rc == 0 ? 0 : (rc == 3 ? 3 : 1));  // (no teardown: the process ends here and the driver reclaims the device)
            }
            close(up[1]);
            Child c;
            c.pid = pid; c.up = up[0];
            children.push_back(c);
        }
        see_only_device(0);
        worker(0);
        int bad = ranks[0].rc;
        for (size_t i = 0; i < children.size(); ++i) {
            int32_t rc = 1;
            if (read(children[i].up, &rc, sizeof(rc)) != (ssize_t)sizeof(rc)) { rc = 1; std::fprintf(stderr, "gpu %zu: worker process died\n", i + 1); }
            close(children[i].up);
            if (rc != 0 && bad == 0) bad = rc;
        }
        if (bad != 0) {
            for (auto& c : children) { int st = 0; waitpid(c.pid, &st, 0); }
            return bad == 3 ? 3 : 1;
        }
        lap("all ranks finished");
        std::vector<void*> staged;
        for (int r = 1; r < gpus; ++r) {
            void* d = nullptr;
            RTX(rtx_malloc(ranks[0].ctx, frame_bytes, &d));
            RTX(rtx_memcpy_h2d(ranks[0].ctx, d, (char*)shared + frame_bytes * (size_t)(r - 1), frame_bytes));
            staged.push_back(d);
            peers.push_back((const float*)d);
        }
        std::vector<uint8_t> rgba_mp((size_t)width * height * 4);
        uint8_t* d_rgba_mp = nullptr;
        RTX(rtx_malloc(ranks[0].ctx, rgba_mp.size(), (void**)&d_rgba_mp));
        RTX(rtx_reduce_tonemap_peers(ranks[0].ctx, ranks[0].accum, peers.data(), (int)peers.size(), width, height, d_rgba_mp));
        RTX(rtx_memcpy_d2h(ranks[0].ctx, rgba_mp.data(), d_rgba_mp, rgba_mp.size()));
        lap("frame on the host");
        RTX(rtx_png_write_rgba8(out.c_str(), width, height, rgba_mp.data()));
        lap("png written");
        double secs_mp = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("%.6fs\n", secs_mp);  // println!("{:?}", instant.elapsed())
        std::fflush(stdout);
        for (auto& c : children) { int st = 0; waitpid(c.pid, &st, 0); }
        _exit(0);  // like the children: no context teardown at the end of the process
    }
    std::vector<std::thread> th;
    for (int r = 0; r < gpus; ++r) th.emplace_back(worker, r);
    for (auto& t : th) t.join();
    for (auto& k : ranks) if (k.rc != 0) return k.rc == 3 ? 3 : 1;

    // combine on rank 0: peers' accumulators are read over NVLink by the fused reduce+tonemap kernel
    std::vector<uint8_t> rgba((size_t)width * height * 4);
    uint8_t* d_rgba = nullptr;
    RTX(rtx_malloc(ranks[0].ctx, rgba.size(), (void**)&d_rgba));
    for (int r = 1; r < gpus; ++r) peers.push_back(ranks[(size_t)r].accum);
    RTX(rtx_reduce_tonemap_peers(ranks[0].ctx, ranks[0].accum, peers.empty() ? nullptr : peers.data(), (int)peers.size(), width, height, d_rgba));
    RTX(rtx_memcpy_d2h(ranks[0].ctx, rgba.data(), d_rgba, rgba.size()));
    lap("frame on the host");
    RTX(rtx_png_write_rgba8(out.c_str(), width, height, rgba.data()));  // image::save_buffer("image.png", ..), main.rs:231
    lap("png written");
    rtx_free(ranks[0].ctx, d_rgba);
    for (auto& k : ranks) {
        rtx_free(k.ctx, k.accum);
        rtx_scene_destroy(k.scene);
        rtx_ctx_destroy(k.ctx);
    }
    rtx_scene_desc_free(desc);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("%.6fs\n", secs);  // println!("{:?}", instant.elapsed())
    return 0;
}
