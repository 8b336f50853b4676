// raytracer.h -- interface-compatible RayTracer (reference include/raytracer.h:15-101).
//
// Same public surface: RayTracer(camera, light), setScene(const Octree*),
// run(w,h), running()/stop()/start(), getImage(); copyable (Gui and Viewer take
// it by value, gui.h:19, viewer.h:16). run() is where the reference loops over
// pixels on one CPU thread; here it hands the frame to lib2019global_b200
// (g19_render: raygen + octree traversal + intersection + shading on the GPU)
// and copies the RGB888 result into a fresh Image. There is no CPU fallback: a
// missing device makes run() throw.
//
// Additions (supersets): setPathTracing(spp, depth, seed) switches run() to the
// wavefront path tracer (G19_MODE_PATH); lastStats() exposes the engine counters.
#pragma once
#include <atomic>
#include <memory>
#include <vector>
#include "camera.h"
#include "entities.h"
#include "image.h"
#include "octree.h"

class RayTracer {
  public:
    RayTracer() = delete;
    RayTracer(const Camera& camera, glm::dvec3 light)
        : _camera(camera), _light(light), _image(std::make_shared<Image>(0, 0)), _engine(std::make_shared<Engine>()) {}
    RayTracer(const RayTracer& o)
        : _running(o._running.load()), _scene(o._scene), _camera(o._camera), _light(o._light), _image(o._image),
          _engine(o._engine), _mode(o._mode), _spp(o._spp), _depth(o._depth), _seed(o._seed), _refresh_ms(o._refresh_ms) {}

    void setScene(const Octree* scene) { _scene = scene; }

    void run(int w, int h) {
        _image = std::make_shared<Image>(w, h);
        if (!_running || !_scene || w <= 0 || h <= 0) return; // nothing renders before start() (raytracer.h:32)
        Engine& e = *_engine;
        std::lock_guard<std::mutex> lock(e.mutex);
        if (!e.ctx) {
            g19_ctx* c = nullptr;
            if (g19_create(nullptr, 0, &c) != G19_OK) throw std::runtime_error(std::string("g19_create: ") + g19_last_error(nullptr));
            e.ctx.reset(c, g19_destroy);
        }
        if (e.scene != _scene->handle() || e.version != _scene->version()) {
            g19::detail::check(g19_upload_scene(e.ctx.get(), _scene->handle()), e.ctx.get(), "g19_upload_scene");
            e.scene = _scene->handle();
            e.version = _scene->version();
        }
        g19_camera cam;
        cam.pos[0] = _camera.pos.x; cam.pos[1] = _camera.pos.y; cam.pos[2] = _camera.pos.z;
        cam.look_at[0] = _camera.lookAtPoint.x; cam.look_at[1] = _camera.lookAtPoint.y; cam.look_at[2] = _camera.lookAtPoint.z;
        cam.focal = _camera.focalDist;
        const double light[3] = {_light.x, _light.y, _light.z};
        g19_params p = {};
        p.width = w; p.height = h; p.mode = _mode; p.spp = _spp; p.max_depth = _depth; p.seed = _seed;
        p.rank = 0; p.world = 1;
        std::vector<uint8_t> rgb(size_t(w) * size_t(h) * 3);
        // incremental rendering (reference raytracer.h:31): the viewer repaints from getImage() every
        // 32 ms (viewer.h:18-21); each refresh copies the samples-so-far image into the live Image
        struct Live { Image* image; int refreshes; } live = {_image.get(), 0};
        int rc = g19_render_progressive(e.ctx.get(), &cam, light, &p, rgb.data(), nullptr,
                                        [](void* u, double, const uint8_t* px) -> int {
                                            Live* l = static_cast<Live*>(u);
                                            l->image->setRows(px);
                                            ++l->refreshes;
                                            return 0;
                                        }, &live, _refresh_ms);
        if (rc != G19_OK && rc != G19_ERR_CANCELLED) g19::detail::check(rc, e.ctx.get(), "g19_render_progressive");
        e.refreshes = live.refreshes;
        g19_get_stats(e.ctx.get(), &e.stats);
    }

    bool running() const { return _running; }
    void stop() {
        _running = false;
        if (_engine->ctx) g19_cancel(_engine->ctx.get()); // callable from the GUI thread while run() is in flight
    }
    void start() { _running = true; }

    std::shared_ptr<Image> getImage() const { return _image; }

    // additions
    void setPathTracing(int spp, int max_depth, unsigned seed = 0) { _mode = G19_MODE_PATH; _spp = spp; _depth = max_depth; _seed = seed; }
    void setReferenceMode() { _mode = G19_MODE_REF; }
    g19_stats lastStats() const { return _engine->stats; }
    void setRefreshInterval(int ms) { _refresh_ms = ms; } // progressive refresh period (default: the viewer's 32 ms timer)
    int lastRefreshes() const { return _engine->refreshes; }

  private:
    struct Engine {
        std::mutex mutex;
        std::shared_ptr<g19_ctx> ctx;
        const g19_scene* scene = nullptr;
        unsigned version = 0;
        g19_stats stats = {};
        int refreshes = 0;
    };
    std::atomic<bool> _running{false};
    const Octree* _scene = nullptr;
    Camera _camera;
    glm::dvec3 _light;
    std::shared_ptr<Image> _image;
    std::shared_ptr<Engine> _engine;
    int _mode = G19_MODE_REF, _spp = 1, _depth = 1;
    unsigned _seed = 0;
    int _refresh_ms = 32;
};
