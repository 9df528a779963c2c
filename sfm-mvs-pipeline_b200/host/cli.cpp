// sfm_match_cli — host driver of the matching stage with the reference's switch grammar
// (-P<key>=<value> parameters, --<flag> flags; AppArgs.cpp:29-53) and switch names
// (PhotogrammetrieCli.cpp:320-392, :95, :112, :422-460):
//   -Pfeature-detector=SIFT|ORB  -Pfeature-matcher=BF|FLANN  -Pfeature-limit=N  -Pfeature-sequence=S
//   -Pfeature-gridlength=L  -Pmatch-threshold=T  --distinct-matches  -Ploglevel=0..4
// Input, either of
//   -Pimage=<shot.pgm|shot.ppm> (repeated, one per shot; binary PGM "P5" or PPM "P6", 8 bit): SfM::extractFeatures runs on the device
//                               (cv::SIFT::create(feature-limit, 3, 0.09), PhotogrammetrieCli.cpp:342-357), the descriptors and
//                               keypoints stay there for the matching and homography stages
// or descriptors computed elsewhere:
//   -Pdescriptors=<bank.sfmd>   "SFMD" u32 version, u32 n_images, u32 cols, u32 depth(0=CV_8U,5=CV_32F),
//                               then per image: u32 n_rows + n_rows*cols*elemsize bytes
//   -Pkeypoints=<bank.sfmk>     optional, enables the homography stage (SfM::calculateHomography): "SFMK" u32 version, u32 n_images,
//                               then per image: u32 n_rows, u32 width, u32 height, n_rows x (float x, float y)
//   -Pransac-matching-threshold=0.006   as PhotogrammetrieCli.cpp:98-99 (< 0: pixels, > 0: fraction of the image size)
//   -Pout-matches-dir=<dir>     with -Pimage: one picture per kept pair (both shots side by side, a line per match), the stage's
//                               on-disk artifact in the reference (PhotogrammetrieCli.cpp:174-199, written there as .jpg; here .ppm)
//   -Pout=<matches.bin>         u64 n_pairs, then per kept pair: i32 left, i32 right, u64 n, n x DMatch(16 B)
//   -Pdevice=<gpu>   or   -Pdevices=<gpu>,<gpu>,...  (one process, one worker thread per GPU; lists gathered on the first)
#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <string>

#include "draw_matches.h"
#include "matching.h"

using namespace sfmhost;

struct Args {
    std::multimap<std::string, std::string> kv;
    void parse(int argc, char** argv) {
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i];
            if (a.size() < 3 || a[0] != '-') { kv.insert({"", a}); continue; }
            const auto eq = a.find('=');
            const std::string type = a.substr(0, 2);
            const std::string key = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            const std::string val = eq == std::string::npos ? "" : a.substr(eq + 1);
            if (type == "-P") kv.insert({key, val});
            else if (type == "--") kv.insert({key, "1"});
        }
    }
    std::string get(const std::string& k, const std::string& def = "") const {
        auto it = kv.find(k);            // first occurrence wins, like AppArgs::getArg
        return it == kv.end() ? def : it->second;
    }
    bool flag(const std::string& k) const { return get(k, "0") == "1"; }
    std::vector<std::string> all(const std::string& k) const {
        std::vector<std::string> v;
        auto r = kv.equal_range(k);
        for (auto it = r.first; it != r.second; ++it) v.push_back(it->second);
        return v;
    }
};

// binary PGM (P5) or PPM (P6), maxval <= 255: what Shot::loadImage hands to the detector.  A colour image goes through the
// grey conversion cv::SIFT applies first (sfm_gray_from_bgr; PPM stores R, G, B).
static void readPnm(const std::string& path, std::vector<uint8_t>& pixels, int& width, int& height) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + path);
    auto token = [&]() {
        std::string t;
        int ch;
        while ((ch = f.get()) != EOF) {
            if (ch == '#') { while ((ch = f.get()) != EOF && ch != '\n') {} continue; }
            if (std::isspace(ch)) { if (!t.empty()) break; continue; }
            t.push_back(static_cast<char>(ch));
        }
        return t;
    };
    const std::string magic = token();
    if (magic != "P5" && magic != "P6") throw std::runtime_error(path + ": not a binary PGM (P5) or PPM (P6)");
    const int channels = magic == "P6" ? 3 : 1;
    width = std::stoi(token());
    height = std::stoi(token());
    const int maxval = std::stoi(token());
    if (width <= 0 || height <= 0 || maxval <= 0 || maxval > 255) throw std::runtime_error(path + ": unsupported PNM header");
    std::vector<uint8_t> raw(static_cast<size_t>(width) * height * channels);
    f.read(reinterpret_cast<char*>(raw.data()), static_cast<std::streamsize>(raw.size()));
    if (!f) throw std::runtime_error(path + ": truncated PNM");
    if (channels == 1) { pixels.swap(raw); return; }
    pixels.resize(static_cast<size_t>(width) * height);
    if (sfm_gray_from_bgr(raw.data(), height, width, 0, 3, 1, pixels.data(), 0) != SFM_OK) throw std::runtime_error(path + ": grey conversion failed");
}

static void usage() {
    std::puts("sfm_match_cli (-Pimage=<shot.pgm> ... | -Pdescriptors=<bank.sfmd>) [-Pfeature-detector=SIFT|ORB] [-Pfeature-matcher=BF|FLANN]\n"
              "              [-Pfeature-limit=10000] [-Pfeature-sequence=0] [-Pfeature-gridlength=0] [-Pmatch-threshold=20]\n"
              "              [--distinct-matches] [-Pkeypoints=<bank.sfmk>] [-Pransac-matching-threshold=0.006]\n"
              "              [-Pout=matches.bin] [-Pdevice=0 | -Pdevices=0,1,...] [-Ploglevel=2]");
}

static void report(const std::vector<ShotMatches>& res, size_t n_pairs, double dt, bool ratios) {
    size_t total = 0;
    for (auto& sm : res) total += sm.matches.size();
    std::printf("pairs=%zu kept=%zu matches=%zu seconds=%.6f\n", n_pairs, res.size(), total, dt);
    if (ratios)
        for (auto& sm : res)
            std::printf("%s : %s -> %zu homographyInlierRatio: %.6f\n", sm.left->imagePath.c_str(), sm.right->imagePath.c_str(),
                        sm.matches.size(), sm.homographyInlierRatio);
}

// -Pimage=...: extractFeatures -> calculateShotMatches -> calculateHomography, all on the device (SfM.cpp:152-156 order)
static int runFromImages(const Args& args, const std::vector<std::string>& paths, const std::string& det, int limit, int loglevel) {
    std::vector<std::vector<uint8_t>> pixels(paths.size());
    std::vector<GrayImage> images(paths.size());
    Scene scene;
    for (size_t i = 0; i < paths.size(); ++i) {
        int w = 0, h = 0;
        readPnm(paths[i], pixels[i], w, h);
        images[i] = GrayImage{pixels[i].data(), h, w, static_cast<size_t>(w)};
        auto shot = std::make_shared<Shot>();
        shot->imagePath = paths[i];
        scene.shots.push_back(shot);
    }
    std::vector<std::string> warnings;
    const bool orb = det == "ORB";                    // PhotogrammetrieCli.cpp:344-356: ORB, else SIFT (+ warning for anything else)
    if (!orb && det != "SIFT" && !det.empty()) warnings.push_back("Unbekannter Merkmalsalgorithmus: " + det + ". Benutze SIFT.");
    std::vector<int> devices;                        // -Pdevices=0,1,...: extraction and matching split over the GPUs of this process
    {
        const std::string dl = args.get("devices");
        size_t pos = 0;
        while (pos < dl.size()) {
            const size_t comma = dl.find(',', pos);
            const std::string tok = dl.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
            if (!tok.empty()) devices.push_back(std::stoi(tok));
            if (comma == std::string::npos) break;
            pos = comma + 1;
        }
        if (devices.empty()) devices.push_back(std::stoi(args.get("device", "0")));
    }
    auto matcher = configureFeatureMatcher(orb ? "ORB" : "SIFT", args.get("feature-matcher"), devices, &warnings);
    auto strategy = configureFeatureMatcherStrategy(std::stoi(args.get("feature-sequence", "0")),
                                                    std::stoi(args.get("feature-gridlength", "0")), &warnings);
    for (auto& w : warnings) std::fprintf(stderr, "[WARN] %s\n", w.c_str());
    std::vector<Features> features;
    const auto t0 = std::chrono::steady_clock::now();
    if (orb) {
        GpuOrbFeatureDetector detector(matcher, limit);               // cv::ORB::create(featureLimit)
        detector.extractFeatures(images, scene, features);
    } else {
        GpuSiftFeatureDetector detector(matcher, limit, 3, 0.09);     // cv::SIFT::create(featureLimit, 3, 0.09)
        detector.extractFeatures(images, scene, features);
    }
    const double t_extract = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    size_t n_kp = 0;
    for (auto& f : features) n_kp += f.keypoints.size();
    std::printf("images=%zu keypoints=%zu extract_seconds=%.6f\n", paths.size(), n_kp, t_extract);
    if (loglevel >= 3)
        for (size_t i = 0; i < paths.size(); ++i) std::printf("%s : %zu Merkmale\n", paths[i].c_str(), features[i].keypoints.size());
    MatchingStage stage;
    stage.setMatchingAlgorithm(matcher);
    stage.setFeatureMatchingStrategy(strategy);
    stage.setMinMatchCount(std::stoi(args.get("match-threshold", "20")));
    stage.setUseDistinctFeatureMatchTest(args.flag("distinct-matches"));
    stage.setRansacReprojectionMatchingThreshold(std::stod(args.get("ransac-matching-threshold", "0.006")));
    const auto t1 = std::chrono::steady_clock::now();
    std::vector<ShotMatches> res = stage.calculateShotMatches(scene);
    stage.calculateHomography(res);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    report(res, strategy->matchPairs(scene.shots.size()).size(), dt, true);
    // the stage's on-disk artifact (PhotogrammetrieCli.cpp:174-199): one picture per ShotMatches, both shots side by side with
    // a line per match, <dir>/<i><left>-<right>.ppm (the reference writes .jpg through cv::imwrite)
    const std::string mdir = args.get("out-matches-dir");
    if (!mdir.empty()) {
        auto base = [](const std::string& p) { const size_t k = p.find_last_of('/'); return k == std::string::npos ? p : p.substr(k + 1); };
        for (size_t i = 0; i < res.size(); ++i) {
            size_t l = 0, r = 0;
            for (size_t k = 0; k < scene.shots.size(); ++k) { if (scene.shots[k] == res[i].left) l = k; if (scene.shots[k] == res[i].right) r = k; }
            const RgbImage img = draw_matches(images[l].data, images[l].rows, images[l].cols, images[l].step, 1,
                                              features[l].keypoints.data(), sizeof(sfm_keypoint),
                                              images[r].data, images[r].rows, images[r].cols, images[r].step, 1,
                                              features[r].keypoints.data(), sizeof(sfm_keypoint),
                                              res[i].matches.data(), res[i].matches.size());
            const std::string path = mdir + "/" + std::to_string(i) + base(paths[l]) + "-" + base(paths[r]) + ".ppm";
            if (!write_ppm(path, img)) throw std::runtime_error("cannot write " + path);
        }
        std::printf("match_pictures=%zu dir=%s\n", res.size(), mdir.c_str());
    }
    return 0;
}

int main(int argc, char** argv) {
    Args args;
    args.parse(argc, argv);
    const std::string path = args.get("descriptors");
    const std::vector<std::string> imagePaths = args.all("image");
    if (path.empty() && imagePaths.empty()) { usage(); return 0; }
    try {
        const int loglevel = std::stoi(args.get("loglevel", "2"));
        const std::string det = args.get("feature-detector");
        int limit = std::stoi(args.get("feature-limit", "10000"));
        if (limit >= SFM_MAX_ROWS) throw std::invalid_argument("feature-limit must stay below 262144");
        if (!imagePaths.empty()) return runFromImages(args, imagePaths, det, limit, loglevel);
        std::ifstream f(path, std::ios::binary);
        if (!f) throw std::runtime_error("cannot open " + path);
        char magic[4]; uint32_t hdr[4];
        f.read(magic, 4); f.read(reinterpret_cast<char*>(hdr), 16);
        if (!f || std::memcmp(magic, "SFMD", 4) != 0 || hdr[0] != 1) throw std::runtime_error("not an SFMD v1 file");
        const uint32_t n_images = hdr[1], cols = hdr[2], depth = hdr[3];
        const size_t esz = depth == SFM_CV_32F ? 4 : 1;
        std::vector<std::vector<char>> storage(n_images);
        Scene scene;
        for (uint32_t i = 0; i < n_images; ++i) {
            uint32_t n = 0;
            f.read(reinterpret_cast<char*>(&n), 4);
            storage[i].resize(static_cast<size_t>(n) * cols * esz);
            f.read(storage[i].data(), static_cast<std::streamsize>(storage[i].size()));
            if (!f) throw std::runtime_error("truncated SFMD file");
            auto shot = std::make_shared<Shot>();
            shot->imagePath = "image" + std::to_string(i);
            // -Pfeature-limit bounds the rows per image (ORB::create(limit) / SIFT::create(limit,...)); 0 = unlimited
            const uint32_t use = (limit > 0 && n > static_cast<uint32_t>(limit)) ? static_cast<uint32_t>(limit) : n;
            shot->descriptors = DescriptorMat{storage[i].data(), static_cast<int>(use), static_cast<int>(cols), cols * esz,
                                              static_cast<int>(depth)};
            scene.shots.push_back(shot);
        }
        const std::string kpath = args.get("keypoints");
        std::vector<std::vector<float>> kstorage(n_images);
        if (!kpath.empty()) {
            std::ifstream kf(kpath, std::ios::binary);
            if (!kf) throw std::runtime_error("cannot open " + kpath);
            char kmagic[4]; uint32_t khdr[2];
            kf.read(kmagic, 4); kf.read(reinterpret_cast<char*>(khdr), 8);
            if (!kf || std::memcmp(kmagic, "SFMK", 4) != 0 || khdr[0] != 1 || khdr[1] != n_images)
                throw std::runtime_error("not an SFMK v1 file for this descriptor bank");
            for (uint32_t i = 0; i < n_images; ++i) {
                uint32_t meta[3];
                kf.read(reinterpret_cast<char*>(meta), 12);
                kstorage[i].resize(static_cast<size_t>(meta[0]) * 2);
                kf.read(reinterpret_cast<char*>(kstorage[i].data()), static_cast<std::streamsize>(kstorage[i].size() * 4));
                if (!kf || static_cast<int>(meta[0]) < scene.shots[i]->descriptors.rows) throw std::runtime_error("truncated SFMK file");
                scene.shots[i]->keypointPts = kstorage[i].data();
                scene.shots[i]->keypointStep = 8;
                scene.shots[i]->imageWidth = static_cast<int>(meta[1]);
                scene.shots[i]->imageHeight = static_cast<int>(meta[2]);
            }
        }
        std::vector<std::string> warnings;
        if (det != "ORB" && det != "SIFT" && !det.empty())
            warnings.push_back("Unbekannter Merkmalsalgorithmus: " + det + ". Benutze SIFT.");
        // -Pdevices=0,1,2,...: several GPUs from this one process (the reference is one process with an OpenMP loop over pairs)
        std::vector<int> devices;
        {
            const std::string dl = args.get("devices");
            size_t pos = 0;
            while (pos < dl.size()) {
                const size_t comma = dl.find(',', pos);
                const std::string tok = dl.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
                if (!tok.empty()) devices.push_back(std::stoi(tok));
                if (comma == std::string::npos) break;
                pos = comma + 1;
            }
            if (devices.empty()) devices.push_back(std::stoi(args.get("device", "0")));
        }
        auto matcher = configureFeatureMatcher(det, args.get("feature-matcher"), devices, &warnings);
        auto strategy = configureFeatureMatcherStrategy(std::stoi(args.get("feature-sequence", "0")),
                                                        std::stoi(args.get("feature-gridlength", "0")), &warnings);
        for (auto& w : warnings) std::fprintf(stderr, "[WARN] %s\n", w.c_str());
        if (matcher->flannMode() && loglevel >= 2)
            std::fprintf(stderr, "[INFO] feature-matcher=FLANN: served by the exact GPU matcher\n");
        MatchingStage stage;
        stage.setMatchingAlgorithm(matcher);
        stage.setFeatureMatchingStrategy(strategy);
        stage.setMinMatchCount(std::stoi(args.get("match-threshold", "20")));
        stage.setUseDistinctFeatureMatchTest(args.flag("distinct-matches"));
        stage.setRansacReprojectionMatchingThreshold(std::stod(args.get("ransac-matching-threshold", "0.006")));
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<ShotMatches> res = stage.calculateShotMatches(scene);
        if (!kpath.empty()) stage.calculateHomography(res);
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const size_t n_pairs = strategy->matchPairs(scene.shots.size()).size();
        size_t total = 0;
        for (auto& sm : res) total += sm.matches.size();
        std::printf("pairs=%zu kept=%zu matches=%zu seconds=%.6f\n", n_pairs, res.size(), total, dt);
        if (loglevel >= 3 || !kpath.empty())
            for (auto& sm : res)
                std::printf("%s : %s -> %zu homographyInlierRatio: %.6f\n", sm.left->imagePath.c_str(), sm.right->imagePath.c_str(),
                            sm.matches.size(), sm.homographyInlierRatio);
        const std::string out = args.get("out");
        if (!out.empty()) {
            std::ofstream o(out, std::ios::binary);
            const uint64_t np = res.size();
            o.write(reinterpret_cast<const char*>(&np), 8);
            for (auto& sm : res) {
                int32_t l = -1, r = -1;
                for (size_t i = 0; i < scene.shots.size(); ++i) {
                    if (scene.shots[i] == sm.left) l = static_cast<int32_t>(i);
                    if (scene.shots[i] == sm.right) r = static_cast<int32_t>(i);
                }
                const uint64_t n = sm.matches.size();
                o.write(reinterpret_cast<const char*>(&l), 4);
                o.write(reinterpret_cast<const char*>(&r), 4);
                o.write(reinterpret_cast<const char*>(&n), 8);
                o.write(reinterpret_cast<const char*>(sm.matches.data()), static_cast<std::streamsize>(n * sizeof(DMatch)));
            }
        }
    } catch (const std::exception& e) {
        // the reference prints the exception and returns 0 (main.cpp:26-34)
        std::fprintf(stderr, "[ERROR] %s\n", e.what());
        return 0;
    }
    return 0;
}
