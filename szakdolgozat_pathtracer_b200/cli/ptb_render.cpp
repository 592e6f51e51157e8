// ptb_render -- command-line front end over the C ABI: the non-interactive render path of the
// reference executable (optixSphere.cpp:754-791 option parsing, 1320-1356 second pass, 1443-1496 --file branch).
//
// Reference options kept with their meaning:
//   --file | -f <filename>     write the image instead of opening a window (the only mode here: there is no display)
//   --dim=<width>x<height>     image size (reference default 1600x1200 in release builds, optixSphere.cpp:759-765)
//   --launch-samples | -s <n>  parsed by the reference into samples_per_launch and then never used
//                              (optixSphere.cpp:1289, 1338-1347); here it sets the samples per pixel per launch
//   --no-gl-interop            accepted and ignored (optixSphere.cpp:1333-1337)
//   --help | -h                usage, exit code 1 (printUsageAndExit, optixSphere.cpp:124-131)
// Everything the reference hard-codes becomes an option with the reference's value as default:
//   --scene a.obj[,b.obj...]   (suitcase.obj,test.obj; optixSphere.cpp:831-834)   --scale 0.05 (optixSphere.cpp:841)
//   --env file.exr             (env4.exr; optixSphere.cpp:835)                     --depth 20 (optixSphere.cu:360)
//   --launches n               number of subframes; the reference's --file branch renders exactly one
//   --seed n                   material seed (the reference uses std::random_device)   --no-dof   --device n
// Beyond the reference:
//   --gpus n                   render on devices `--device` .. `--device` + n - 1 of this box (ptb_multi, one process)
//   --split samples|tiles      how a launch is divided between the GPUs (default samples; tiles is bit-identical to one GPU)
//   --batch k                  subframes per launch (one wavefront of k * width * height paths; with --gpus the unit that is split)
//   --fast                     fast arithmetic (ptb_render_cfg.arith_mode = 1)
//   --accum-out file           also write the float4 accumulation buffer (ptb_save_accum_raw)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ptb.h"

static void usage_and_exit(const char* argv0) {
    fprintf(stderr, "Usage  : %s [options]\n", argv0);
    fprintf(stderr, "Options: --file | -f <filename>      Specify file for image output (.png or .ppm)\n");
    fprintf(stderr, "         --help | -h                 Print this usage message\n");
    fprintf(stderr, "         --dim=<width>x<height>      Set image dimensions; defaults to 1600x1200\n");
    fprintf(stderr, "         --launch-samples | -s <n>   Samples per pixel per launch (default 10)\n");
    fprintf(stderr, "         --scene a.obj[,b.obj]  --env file.exr  --scale s  --depth d  --launches n  --seed n  --no-dof  --device n\n");
    fprintf(stderr, "         --gpus n  --split samples|tiles  --batch k  --fast  --accum-out file\n");
    exit(1);
}

#define CHECK(call) do { if ((call) != PTB_OK) { fprintf(stderr, "Caught exception: %s\n", ptb_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
    std::string outfile, env = "env4.exr", scene_arg = "suitcase.obj,test.obj";
    int width = 1600, height = 1200, spp = 10, depth = 20, launches = 1, device = 0, gpus = 1, batch = 1, split = PTB_SPLIT_SAMPLES;
    bool fast = false;
    std::string accum_out;
    unsigned seed = 1;
    float scale = 0.05f;
    bool dof = true;
    for (int i = 1; i < argc; ++i) {
        const std::string arg(argv[i]);
        auto next = [&](const char* what) -> const char* { if (i >= argc - 1) { fprintf(stderr, "Option '%s' needs a value\n", what); usage_and_exit(argv[0]); } return argv[++i]; };
        if (arg == "--help" || arg == "-h") usage_and_exit(argv[0]);
        else if (arg == "--file" || arg == "-f") outfile = next("--file");
        else if (arg.substr(0, 6) == "--dim=") {
            if (sscanf(arg.c_str() + 6, "%dx%d", &width, &height) != 2 || width <= 0 || height <= 0) { fprintf(stderr, "Invalid window dimensions '%s'\n", arg.c_str() + 6); usage_and_exit(argv[0]); }
        } else if (arg == "--launch-samples" || arg == "-s") spp = atoi(next("--launch-samples"));
        else if (arg == "--no-gl-interop") {}
        else if (arg == "--scene") scene_arg = next("--scene");
        else if (arg == "--env") env = next("--env");
        else if (arg == "--scale") scale = (float)atof(next("--scale"));
        else if (arg == "--depth") depth = atoi(next("--depth"));
        else if (arg == "--launches") launches = atoi(next("--launches"));
        else if (arg == "--seed") seed = (unsigned)strtoul(next("--seed"), nullptr, 10);
        else if (arg == "--device") device = atoi(next("--device"));
        else if (arg == "--no-dof") dof = false;
        else if (arg == "--gpus") gpus = atoi(next("--gpus"));
        else if (arg == "--batch") batch = atoi(next("--batch"));
        else if (arg == "--fast") fast = true;
        else if (arg == "--accum-out") accum_out = next("--accum-out");
        else if (arg == "--split") {
            const std::string v = next("--split");
            if (v == "samples") split = PTB_SPLIT_SAMPLES; else if (v == "tiles") split = PTB_SPLIT_TILES;
            else { fprintf(stderr, "Unknown split '%s'\n", v.c_str()); usage_and_exit(argv[0]); }
        }
        else { fprintf(stderr, "Unknown option '%s'\n", argv[i]); usage_and_exit(argv[0]); }
    }
    if (outfile.empty()) { fprintf(stderr, "No display is available on this platform: --file <filename> is required\n"); usage_and_exit(argv[0]); }
    if (spp < 1 || depth < 0 || launches < 1 || gpus < 1 || gpus > 16 || batch < 1) usage_and_exit(argv[0]);

    std::vector<std::string> files;
    for (size_t p = 0; p <= scene_arg.size();) {
        size_t q = scene_arg.find(',', p);
        if (q == std::string::npos) q = scene_arg.size();
        if (q > p) files.push_back(scene_arg.substr(p, q - p));
        p = q + 1;
    }
    std::vector<const char*> cfiles;
    for (const std::string& f : files) cfiles.push_back(f.c_str());

    // one process, `gpus` devices; device index 0 of the group is the root that owns the accumulator and the output buffer
    std::vector<int> devices;
    for (int g = 0; g < gpus; ++g) devices.push_back(device + g);
    ptb_multi* multi = nullptr;
    CHECK(ptb_multi_create(devices.data(), gpus, &multi));
    ptb_context* ctx = ptb_multi_context(multi, 0);
    ptb_scene* scene = nullptr;
    CHECK(ptb_scene_load_obj(cfiles.data(), (int)cfiles.size(), scale, seed, &scene));
    printf("Loaded models with %u triangles total.\n", ptb_scene_num_triangles(scene));
    CHECK(ptb_scene_set_env_file(scene, env.c_str()));

    ptb_Params params;
    memset(&params, 0, sizeof(params));
    params.image_width = (unsigned)width; params.image_height = (unsigned)height;
    params.origin_x = width / 2; params.origin_y = height / 2;
    params.subframe_index = 0; params.dof = dof;
    ptb_params_default_camera(&params);
    ptb_build_stats bst;
    CHECK(ptb_multi_accel_build(multi, scene, nullptr, &bst));
    printf("BVH: %u nodes, depth %u, SAH %.2f, %.3f ms, built on %d GPU%s\n", bst.num_nodes, bst.max_depth, bst.sah_cost, bst.build_ms, gpus, gpus > 1 ? "s" : "");
    void* accum = nullptr;
    CHECK(ptb_device_alloc(ctx, (size_t)width * height * sizeof(ptb_float4), &accum));
    CHECK(ptb_device_memset(ctx, accum, 0, (size_t)width * height * sizeof(ptb_float4), nullptr));
    CHECK(ptb_context_synchronize(ctx, nullptr));
    params.accum_buffer = (ptb_float4*)accum;
    ptb_output* out = nullptr;
    CHECK(ptb_output_create(ctx, (unsigned)width, (unsigned)height, &out));
    ptb_render_cfg cfg;
    ptb_default_render_cfg(&cfg);
    cfg.spp_per_launch = spp; cfg.max_depth = depth; cfg.subframes_per_launch = batch; cfg.arith_mode = fast ? PTB_ARITH_FAST : PTB_ARITH_EXACT;
    for (int l = 0; l < launches; ++l) {
        params.frame_buffer = ptb_output_map(out);
        CHECK(ptb_multi_launch(multi, &params, &cfg, split));
        CHECK(ptb_multi_synchronize(multi));
        ptb_output_unmap(out, nullptr);
        params.subframe_index += batch;
    }
    uint64_t totals[4];
    CHECK(ptb_multi_get_totals(multi, totals, 0));
    const unsigned long long segments = totals[0];
    const ptb_uchar4* host = ptb_output_host_ptr(out);
    if (!host) { fprintf(stderr, "Caught exception: %s\n", ptb_last_error()); return 1; }
    CHECK(ptb_save_image(outfile.c_str(), host, width, height, 1));
    if (!accum_out.empty()) {
        std::vector<ptb_float4> h_accum((size_t)width * height);
        CHECK(ptb_copy_to_host(ctx, h_accum.data(), accum, h_accum.size() * sizeof(ptb_float4), nullptr));
        CHECK(ptb_save_accum_raw(accum_out.c_str(), h_accum.data(), width, height));
    }
    printf("Wrote %s (%dx%d, %d spp, %llu segments, %d GPU%s%s)\n", outfile.c_str(), width, height, spp * launches * batch, segments, gpus,
           gpus > 1 ? "s" : "", gpus > 1 ? (split == PTB_SPLIT_TILES ? ", tile split" : ", sample split") : "");
    ptb_output_destroy(out);
    ptb_device_free(ctx, accum);
    ptb_scene_destroy(scene);
    ptb_multi_destroy(multi);
    return 0;
}
