// test_match_index — ShotMatchIndex against the reference's linear scans (restated below from Scene.cpp:327-338, :385-412,
// :504-545) on random scenes: duplicated points, both orientations, -0 / NaN coordinates.  Prints "OK <checks>" or the
// first mismatch.  Built with g++ only (no CUDA): `make -C sfm-mvs-pipeline_b200 test_match_index`.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "match_index.h"

using namespace sfmhost;

struct ShotData { std::vector<float> pts; };                       // x, y per keypoint (packed, step 8)
struct SM { int l, r; std::vector<sfm_dmatch> m; };

static bool ptEq(const float* p, double x, double y) { return static_cast<double>(p[0]) == x && static_cast<double>(p[1]) == y; }

int main(int argc, char** argv) {
    const unsigned seed = argc > 1 ? static_cast<unsigned>(std::atoi(argv[1])) : 1u;
    std::mt19937 rng(seed);
    long checks = 0;
    for (int trial = 0; trial < 40; ++trial) {
        const int n_shots = 2 + static_cast<int>(rng() % 9);
        std::vector<ShotData> shots(n_shots);
        for (auto& s : shots) {
            const int nk = 1 + static_cast<int>(rng() % 60);
            for (int k = 0; k < nk; ++k) {
                // a small coordinate alphabet creates many keypoints with identical pt (cv::SIFT: several orientations)
                float x = static_cast<float>(rng() % 12) * 0.5f, y = static_cast<float>(rng() % 7) * 0.25f;
                if (rng() % 37 == 0) x = -0.0f;
                if (rng() % 53 == 0) y = std::nanf("");
                s.pts.push_back(x); s.pts.push_back(y);
            }
        }
        std::vector<SM> sms;
        const int n_sm = static_cast<int>(rng() % 25);
        for (int i = 0; i < n_sm; ++i) {
            SM e;
            e.l = static_cast<int>(rng() % n_shots); e.r = static_cast<int>(rng() % n_shots);
            const int nm = static_cast<int>(rng() % 80);
            for (int k = 0; k < nm; ++k) {
                sfm_dmatch d;
                d.queryIdx = static_cast<int>(rng() % (shots[e.l].pts.size() / 2));
                d.trainIdx = static_cast<int>(rng() % (shots[e.r].pts.size() / 2));
                d.imgIdx = 0;
                d.distance = static_cast<float>(rng() % 400) - 50.0f;
                e.m.push_back(d);
            }
            sms.push_back(std::move(e));
        }
        // ---- Scene::addShotMatches restated + the index
        std::vector<int> kept;                                     // indices into sms, in insertion order
        ShotMatchIndex index;
        for (int i = 0; i < n_sm; ++i) {
            int existing = -1;
            for (std::size_t k = 0; k < kept.size(); ++k)
                if (sms[kept[k]].l == sms[i].l && sms[kept[k]].r == sms[i].r) { existing = static_cast<int>(k); break; }
            IndexedShotMatches e;
            e.left = &shots[sms[i].l]; e.right = &shots[sms[i].r];
            e.matches = sms[i].m.data(); e.n_matches = sms[i].m.size();
            e.leftPts = shots[sms[i].l].pts.data(); e.rightPts = shots[sms[i].r].pts.data();
            bool inserted = false;
            const int got = index.add(e, &inserted);
            const int want = existing >= 0 ? existing : static_cast<int>(kept.size());
            if (got != want || inserted != (existing < 0)) { std::printf("addShotMatches mismatch trial %d entry %d\n", trial, i); return 1; }
            if (existing < 0) kept.push_back(i);
            ++checks;
        }
        // ---- lookups
        for (int q = 0; q < 400; ++q) {
            const int a = static_cast<int>(rng() % n_shots), b = static_cast<int>(rng() % n_shots);
            int want = -1;
            for (std::size_t k = 0; k < kept.size(); ++k) {
                const SM& e = sms[kept[k]];
                if ((e.l == a && e.r == b) || (e.r == a && e.l == b)) { want = static_cast<int>(k); break; }
            }
            const int got = index.findEither(&shots[a], &shots[b]);
            if (got != want) { std::printf("findEither mismatch trial %d\n", trial); return 1; }
            ++checks;
            if (want < 0) continue;
            const SM& e = sms[kept[want]];
            // find3d2dMatches: keypoint on the side of `origin` equals the point
            const bool originIsLeft = rng() % 2 == 0;
            const ShotData& os = shots[originIsLeft ? e.l : e.r];
            const int kp = static_cast<int>(rng() % (os.pts.size() / 2));
            double px = os.pts[2 * kp], py = os.pts[2 * kp + 1];
            if (rng() % 9 == 0) px += 1e-9;                        // a double that is not a float: equals nothing
            int wm = -1;
            for (std::size_t i = 0; i < e.m.size(); ++i) {
                const float* p = originIsLeft ? &shots[e.l].pts[2 * e.m[i].queryIdx] : &shots[e.r].pts[2 * e.m[i].trainIdx];
                if (ptEq(p, px, py)) { wm = static_cast<int>(i); break; }
            }
            if (index.firstMatchWithPoint(want, originIsLeft, px, py) != wm) { std::printf("firstMatchWithPoint mismatch trial %d\n", trial); return 1; }
            ++checks;
            // mergePointcloudElement3d2d
            const int kl = static_cast<int>(rng() % (shots[e.l].pts.size() / 2)), kr = static_cast<int>(rng() % (shots[e.r].pts.size() / 2));
            const double lx = shots[e.l].pts[2 * kl], ly = shots[e.l].pts[2 * kl + 1];
            const double rx = shots[e.r].pts[2 * kr], ry = shots[e.r].pts[2 * kr + 1];
            const double bound = static_cast<double>(rng() % 300);
            int wm2 = -1;
            for (std::size_t i = 0; i < e.m.size(); ++i)
                if (ptEq(&shots[e.l].pts[2 * e.m[i].queryIdx], lx, ly) && ptEq(&shots[e.r].pts[2 * e.m[i].trainIdx], rx, ry) &&
                    std::fabs(static_cast<double>(e.m[i].distance)) <= bound) { wm2 = static_cast<int>(i); break; }
            if (index.firstMatchWithPoints(want, lx, ly, rx, ry, bound) != wm2) { std::printf("firstMatchWithPoints mismatch trial %d\n", trial); return 1; }
            ++checks;
        }
    }
    std::printf("OK %ld\n", checks);
    return 0;
}
