// g19_headless.cpp -- what reference main.cpp:19-62 + viewer.h:46-61 do, minus Qt:
// build the default scene through the interface headers (include/*.h), start the
// ray tracer on a worker thread, run(500,500), write the frame as binary PPM.
//   g19_headless out.ppm|out.png [width height] [--path spp depth] [--refresh ms]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "camera.h"
#include "entities.h"
#include "octree.h"
#include "raytracer.h"

int main(int argc, char** argv) {
    const char* out = argc > 1 ? argv[1] : "render.ppm";
    int w = 500, h = 500, spp = 0, depth = 0;
    if (argc > 3 && argv[2][0] != '-') { w = std::atoi(argv[2]); h = std::atoi(argv[3]); }
    for (int i = 2; i + 2 < argc; ++i)
        if (!std::strcmp(argv[i], "--path")) { spp = std::atoi(argv[i + 1]); depth = std::atoi(argv[i + 2]); }

    Camera camera({-10, 0, 0}, {1, 0, 0}, 0.1);
    glm::dvec3 light{-10, 10, 10};
    RayTracer raytracer(camera, light);
    Octree scene({-20, -20, -20}, {20, 20, 20});
    ImpSphere* s2 = new ImpSphere(glm::dvec3{3, 4, 4}, 2, {1, 0, 0});
    ImpSphere* s3 = new ImpSphere(glm::dvec3{4, -4, 4}, 2, {0, 0, 1});
    ExpQuad* q = new ExpQuad(glm::dvec3{0, 0, 0}, 2, 3, (90.0 * M_PI / 180.0), {1, 2, 3});
    scene.push_back(q);
    scene.push_back(s2);
    scene.push_back(s3);
    raytracer.setScene(&scene);
    if (spp > 0) raytracer.setPathTracing(spp, depth);
    for (int i = 2; i + 1 < argc; ++i)
        if (!std::strcmp(argv[i], "--refresh")) raytracer.setRefreshInterval(std::atoi(argv[i + 1]));

    RayTracer worker_copy(raytracer); // Gui/Viewer copy the tracer by value
    worker_copy.start();
    double seconds = 0;
    std::thread t([&] {
        auto t1 = std::chrono::high_resolution_clock::now();
        worker_copy.run(w, h);
        seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t1).count();
    });
    t.join();
    std::shared_ptr<Image> img = worker_copy.getImage();
    if (!img->save(out)) return 2; // .png (what the reference's "Save as..." writes, gui.h:39-45) or .ppm
    std::printf("%dx%d in %.4f seconds, %d progressive refreshes -> %s\n", img->width(), img->height(), seconds,
                worker_copy.lastRefreshes(), out);
    // interface check: the host-callable virtuals answer through the GPU probes
    glm::dvec3 p, n;
    Ray r(glm::dvec3{-10, 0, 0}, glm::dvec3{13, 4, 4});
    bool hit = s2->intersect(r, p, n);
    std::printf("ImpSphere::intersect -> %d at (%.6f, %.6f, %.6f); candidates %zu\n", int(hit), p.x, p.y, p.z,
                scene.intersect(r).size());
    return 0;
}
