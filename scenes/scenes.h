// scenes.h — the BASELINE.json scenes, written ONCE against the reference's scene API
// (make_shared<sphere/quad/triangle/...>, box(), mesh::loadObj, translate, rotate_y,
// constant_medium, bvh_node, hittable_list, camera fields).  This file includes no
// renderer header itself: the translation unit that includes it decides which header
// set it compiles against —
//   * oracle/ref_driver.cpp includes the UNTOUCHED reference headers from
//     /root/reference first  -> the scenes feed the reference CPU renderer (the oracle);
//   * scenes/scenes_capi.cpp and apps/ include the host mirror
//     (raytracingoneweekendapplication_b200/host) first -> the same scenes are flattened
//     and rendered by the CUDA library.
// Both builds draw their construction-time random numbers from the same generator
// (host_rng.h; the oracle driver interposes rand() with it), so a given seed yields the
// same world in both.  Random draws are sequenced explicitly (separate statements)
// because argument evaluation order is unspecified in C++.
//
// Scene sources: C2 = main.cpp:208-243, C3 = main.cpp:341-380, C5 = main.cpp:268-340
// (with the shadowing `hittable_list world` of main.cpp:288 removed),
// `quads` = main.cpp:176-192, `emissive` = main.cpp:244-267, `specular` = main.cpp:381-439
// (the default scene 7), `mixed` = main.cpp:128-175 (scene 0).  C1 (book-1 final) is not
// in main.cpp and follows the standard RTOW layout given in SURVEY §8d.
// C4's corgi/car assets are missing from the reference (.MISSING_LARGE_BLOBS), so it
// loads procedurally generated OBJ/PPM assets (raytracingoneweekendapplication_b200/assets.py).
#ifndef RTB200_SCENES_H
#define RTB200_SCENES_H

#include <string>
#include <vector>

struct scene_config {
    std::string asset_dir = "scenes/assets";
    // the BASELINE.json frame for the scene (the caller may override)
    int width = 400, height = 225, spp = 10, depth = 50;
    bool wrap_in_bvh = true;  // main.cpp:442
};

namespace scenes_detail {

inline color rand_color() {
    double r = random_double();
    double g = random_double();
    double b = random_double();
    return color(r, g, b);
}
inline color rand_color(double lo, double hi) {
    double r = random_double(lo, hi);
    double g = random_double(lo, hi);
    double b = random_double(lo, hi);
    return color(r, g, b);
}

inline void set_frame(camera& cam, scene_config& cfg, int w, int h, int spp) {
    cfg.width = w; cfg.height = h; cfg.spp = spp; cfg.depth = 50;
    cam.image_width = w;
    cam.aspect_ratio = double(w) / double(h);
    cam.samples_per_pixel = spp;
    cam.max_depth = 50;
}

// C1: RTOW book-1 final scene, 400x225, 10 spp
inline void book1(hittable_list& world, camera& cam, scene_config& cfg) {
    world.add(make_shared<sphere>(point3(0, -1000, 0), 1000, make_shared<lambertian>(color(0.5, 0.5, 0.5))));
    for (int a = -11; a < 11; a++) {
        for (int b = -11; b < 11; b++) {
            double choose_mat = random_double();
            double jx = random_double();
            double jz = random_double();
            point3 center(a + 0.9 * jx, 0.2, b + 0.9 * jz);
            if ((center - point3(4, 0.2, 0)).length() > 0.9) {
                if (choose_mat < 0.8) {
                    color c1 = rand_color();
                    color c2 = rand_color();
                    world.add(make_shared<sphere>(center, 0.2, make_shared<lambertian>(c1 * c2)));
                } else if (choose_mat < 0.95) {
                    color albedo = rand_color(0.5, 1);
                    double fuzz = random_double(0, 0.5);
                    world.add(make_shared<sphere>(center, 0.2, make_shared<metal>(albedo, fuzz)));
                } else {
                    world.add(make_shared<sphere>(center, 0.2, make_shared<dielectric>(1.5)));
                }
            }
        }
    }
    world.add(make_shared<sphere>(point3(0, 1, 0), 1.0, make_shared<dielectric>(1.5)));
    world.add(make_shared<sphere>(point3(-4, 1, 0), 1.0, make_shared<lambertian>(color(0.4, 0.2, 0.1))));
    world.add(make_shared<sphere>(point3(4, 1, 0), 1.0, make_shared<metal>(color(0.7, 0.6, 0.5), 0.0)));

    set_frame(cam, cfg, 400, 225, 10);
    cam.background = color(0.7, 0.8, 1.0);
    cam.vfov = 20;
    cam.lookfrom = point3(13, 2, 3);
    cam.lookat = point3(0, 0, 0);
    cam.vup = vec3(0, 1, 0);
    cam.defocus_angle = 0.6;
    cam.focus_dist = 10.0;
}

inline void cornell_walls(hittable_list& world, const point3& light_q, const vec3& light_u, const vec3& light_v,
                          double light_power, bool book_ceiling) {
    auto red = make_shared<lambertian>(color(.65, .05, .05));
    auto white = make_shared<lambertian>(color(.73, .73, .73));
    auto green = make_shared<lambertian>(color(.12, .45, .15));
    auto light = make_shared<diffuse_light>(color(light_power, light_power, light_power));
    world.add(make_shared<quad>(point3(555, 0, 0), vec3(0, 555, 0), vec3(0, 0, 555), green));
    world.add(make_shared<quad>(point3(0, 0, 0), vec3(0, 555, 0), vec3(0, 0, 555), red));
    world.add(make_shared<quad>(light_q, light_u, light_v, light));
    if (book_ceiling) {  // main.cpp:218-219 order
        world.add(make_shared<quad>(point3(0, 0, 0), vec3(555, 0, 0), vec3(0, 0, 555), white));
        world.add(make_shared<quad>(point3(555, 555, 555), vec3(-555, 0, 0), vec3(0, 0, -555), white));
    } else {             // main.cpp:353-354 order
        world.add(make_shared<quad>(point3(0, 555, 0), vec3(555, 0, 0), vec3(0, 0, 555), white));
        world.add(make_shared<quad>(point3(0, 0, 0), vec3(555, 0, 0), vec3(0, 0, 555), white));
    }
    world.add(make_shared<quad>(point3(0, 0, 555), vec3(555, 0, 0), vec3(0, 555, 0), white));
}

inline void cornell_camera(camera& cam) {
    cam.background = color(0, 0, 0);
    cam.vfov = 40;
    cam.lookfrom = point3(278, 278, -800);
    cam.lookat = point3(278, 278, 0);
    cam.vup = vec3(0, 1, 0);
    cam.defocus_angle = 0;
}

// C2: Cornell box, 600x600, 200 spp (main.cpp:208-243)
inline void cornell(hittable_list& world, camera& cam, scene_config& cfg) {
    cornell_walls(world, point3(343, 554, 332), vec3(-130, 0, 0), vec3(0, 0, -105), 15, true);
    auto white = make_shared<lambertian>(color(.73, .73, .73));
    shared_ptr<hittable> box1 = box(point3(0, 0, 0), point3(165, 330, 165), white);
    box1 = make_shared<rotate_y>(box1, 15);
    box1 = make_shared<translate>(box1, vec3(265, 0, 295));
    world.add(box1);
    shared_ptr<hittable> box2 = box(point3(0, 0, 0), point3(165, 165, 165), white);
    box2 = make_shared<rotate_y>(box2, -18);
    box2 = make_shared<translate>(box2, vec3(130, 0, 65));
    world.add(box2);
    set_frame(cam, cfg, 600, 600, 200);
    cornell_camera(cam);
}

// C3: Cornell box with two smoke boxes, 600x600, 1000 spp (main.cpp:341-380)
inline void cornell_smoke(hittable_list& world, camera& cam, scene_config& cfg) {
    cornell_walls(world, point3(113, 554, 127), vec3(330, 0, 0), vec3(0, 0, 305), 7, false);
    auto white = make_shared<lambertian>(color(.73, .73, .73));
    shared_ptr<hittable> box1 = box(point3(0, 0, 0), point3(165, 330, 165), white);
    box1 = make_shared<rotate_y>(box1, 15);
    box1 = make_shared<translate>(box1, vec3(265, 0, 295));
    shared_ptr<hittable> box2 = box(point3(0, 0, 0), point3(165, 165, 165), white);
    box2 = make_shared<rotate_y>(box2, -18);
    box2 = make_shared<translate>(box2, vec3(130, 0, 65));
    world.add(make_shared<constant_medium>(box1, 0.005, color(0, 0, 0)));
    world.add(make_shared<constant_medium>(box2, 0.005, color(0.2, 0.2, 0.2)));
    set_frame(cam, cfg, 600, 600, 1000);
    cornell_camera(cam);
}

// C4: triangle-mesh scene, 1920x1080, 64 spp.  Three OBJ meshes through mesh::loadObj
// with glm transforms composed like main.cpp:399-405 (translate * rotate_y * scale),
// an image texture, a UV checker, a Perlin texture, a ground sphere, one point light.
inline void mesh_scene(hittable_list& world, camera& cam, std::vector<point_light>& lights, scene_config& cfg) {
    auto grey = make_shared<lambertian>(color(0.35, 0.35, 0.35));
    world.add(make_shared<sphere>(point3(0, -1000, 0), 1000, grey));

    auto hide_tex = make_shared<image_texture>((cfg.asset_dir + "/earth.ppm").c_str());
    auto hide_mat = make_shared<lambertian>(hide_tex);
    auto checkerT = make_shared<checker_texture_triangle>(0.5, color(0.05, 0.05, 0.05), color(.9, .9, .9));
    auto checker_mat = make_shared<lambertian>(checkerT);
    auto marble = make_shared<lambertian>(make_shared<noise_texture>(4));

    {
        mesh knot;
        glm::mat4 tf = glm::mat4(1.0f);
        tf = glm::translate(tf, glm::vec3(0.0f, 2.2f, 0.0f));
        tf = glm::rotate(tf, glm::radians(90.0f), glm::vec3(0, 1, 0));
        tf = glm::scale(tf, glm::vec3(1.5f, 1.5f, 1.5f));
        knot.loadObj(cfg.asset_dir + "/knot.obj", world, hide_mat, tf);
    }
    {
        mesh blob;
        glm::mat4 tf = glm::mat4(1.0f);
        tf = glm::translate(tf, glm::vec3(-4.2f, 1.4f, 1.0f));
        tf = glm::rotate(tf, glm::radians(35.0f), glm::vec3(0, 1, 0));
        tf = glm::scale(tf, glm::vec3(1.3f, 1.3f, 1.3f));
        blob.loadObj(cfg.asset_dir + "/blob.obj", world, checker_mat, tf);
    }
    {
        mesh crate;
        glm::mat4 tf = glm::mat4(1.0f);
        tf = glm::translate(tf, glm::vec3(4.0f, 1.0f, 0.5f));
        tf = glm::rotate(tf, glm::radians(-25.0f), glm::vec3(0, 1, 0));
        crate.loadObj(cfg.asset_dir + "/crate.obj", world, marble, tf);
    }
    world.add(make_shared<sphere>(point3(2.2, 0.6, -2.6), 0.6, make_shared<dielectric>(1.5)));
    world.add(make_shared<sphere>(point3(-1.8, 0.5, -3.0), 0.5, make_shared<metal>(color(0.8, 0.7, 0.4), 0.1)));

    lights.push_back(point_light(point3(0, 9, -6), color(40, 38, 34), 0.5));

    set_frame(cam, cfg, 1920, 1080, 64);
    cam.background = color(0.25, 0.32, 0.45);
    cam.vfov = 35;
    cam.lookfrom = point3(0, 4.5, -13);
    cam.lookat = point3(0, 1.6, 0);
    cam.vup = vec3(0, 1, 0);
    cam.defocus_angle = 0;
    cam.focus_dist = 10;
}

// C5: RTOW book-2 final scene, 3840x2160, 1024 spp (main.cpp:268-340)
inline void final_scene(hittable_list& world, camera& cam, scene_config& cfg) {
    hittable_list boxes1;
    auto ground = make_shared<lambertian>(color(0.48, 0.83, 0.53));
    const int boxes_per_side = 20;
    for (int i = 0; i < boxes_per_side; i++) {
        for (int j = 0; j < boxes_per_side; j++) {
            double w = 100.0;
            double x0 = -1000.0 + i * w;
            double z0 = -1000.0 + j * w;
            double y0 = 0.0;
            double x1 = x0 + w;
            double y1 = random_double(1, 101);
            double z1 = z0 + w;
            boxes1.add(box(point3(x0, y0, z0), point3(x1, y1, z1), ground));
        }
    }
    world.add(make_shared<bvh_node>(boxes1));

    auto light = make_shared<diffuse_light>(color(7, 7, 7));
    world.add(make_shared<quad>(point3(123, 554, 147), vec3(300, 0, 0), vec3(0, 0, 265), light));

    point3 center1(400, 400, 200);
    point3 center2 = center1 + vec3(30, 0, 0);
    auto sphere_material = make_shared<lambertian>(color(0.7, 0.3, 0.1));
    world.add(make_shared<sphere>(center1, center2, 50, sphere_material));

    world.add(make_shared<sphere>(point3(260, 150, 45), 50, make_shared<dielectric>(1.5)));
    world.add(make_shared<sphere>(point3(0, 150, 145), 50, make_shared<metal>(color(0.8, 0.8, 0.9), 1.0)));

    auto boundary = make_shared<sphere>(point3(360, 150, 145), 70, make_shared<dielectric>(1.5));
    world.add(boundary);
    world.add(make_shared<constant_medium>(boundary, 0.2, color(0.2, 0.4, 0.9)));
    boundary = make_shared<sphere>(point3(0, 0, 0), 5000, make_shared<dielectric>(1.5));
    world.add(make_shared<constant_medium>(boundary, .0001, color(1, 1, 1)));

    auto emat = make_shared<lambertian>(make_shared<image_texture>((cfg.asset_dir + "/earth.ppm").c_str()));
    world.add(make_shared<sphere>(point3(400, 200, 400), 100, emat));
    auto pertext = make_shared<noise_texture>(0.2);
    world.add(make_shared<sphere>(point3(220, 280, 300), 80, make_shared<lambertian>(pertext)));

    hittable_list boxes2;
    auto white = make_shared<lambertian>(color(.73, .73, .73));
    const int ns = 1000;
    for (int j = 0; j < ns; j++) {
        double px = random_double(0, 165);
        double py = random_double(0, 165);
        double pz = random_double(0, 165);
        boxes2.add(make_shared<sphere>(point3(px, py, pz), 10, white));
    }
    world.add(make_shared<translate>(make_shared<rotate_y>(make_shared<bvh_node>(boxes2), 15), vec3(-100, 270, 395)));

    set_frame(cam, cfg, 3840, 2160, 1024);
    cam.background = color(0, 0, 0);
    cam.vfov = 40;
    cam.lookfrom = point3(478, 278, -600);
    cam.lookat = point3(278, 278, 0);
    cam.vup = vec3(0, 1, 0);
    cam.defocus_angle = 0;
}

// main.cpp:176-192 (scene 1): five coloured quads, camera defaults
inline void quads(hittable_list& world, camera& cam, scene_config& cfg) {
    world.add(make_shared<quad>(point3(-3, -2, 5), vec3(0, 0, -4), vec3(0, 4, 0), make_shared<lambertian>(color(1.0, 0.2, 0.2))));
    world.add(make_shared<quad>(point3(-2, -2, 0), vec3(4, 0, 0), vec3(0, 4, 0), make_shared<lambertian>(color(0.2, 1.0, 0.2))));
    world.add(make_shared<quad>(point3(3, -2, 1), vec3(0, 0, 4), vec3(0, 4, 0), make_shared<lambertian>(color(0.2, 0.2, 1.0))));
    world.add(make_shared<quad>(point3(-2, 3, 1), vec3(4, 0, 0), vec3(0, 0, 4), make_shared<lambertian>(color(1.0, 0.5, 0.0))));
    world.add(make_shared<quad>(point3(-2, -3, 5), vec3(4, 0, 0), vec3(0, 0, -4), make_shared<lambertian>(color(0.2, 0.8, 0.8))));
    set_frame(cam, cfg, 400, 400, 100);
    cam.background = color(0.70, 0.80, 1.00);
    cam.vfov = 80;
    cam.lookfrom = point3(0, 0, 9);
    cam.lookat = point3(0, 0, 0);
    cam.vup = vec3(0, 1, 0);
    cam.defocus_angle = 0;
}

// main.cpp:244-267 (scene 4): emissive_light sphere above a red sphere
inline void emissive(hittable_list& world, camera& cam, scene_config& cfg) {
    auto red = make_shared<lambertian>(color(.65, .05, .05));
    world.add(make_shared<sphere>(point3(0, 2, 4), 1.0, red));
    shared_ptr<material> glow = make_shared<emissive_light>(color(1.0, 1.0, 1.0));
    world.add(make_shared<sphere>(point3(0, 4, 0), 3, glow));
    set_frame(cam, cfg, 400, 225, 200);
    cam.max_depth = 5; cfg.depth = 5;
    cam.background = color(0, 0, 0);
    cam.vfov = 40;
    cam.lookfrom = point3(0, 0, 0);
    cam.lookat = point3(0, 2, 4);
    cam.vup = vec3(0, 1, 0);
}

// main.cpp:381-439 (scene 7, the reference's default): the author's `specular` material
inline void specular_scene(hittable_list& world, camera& cam, scene_config& cfg) {
    auto grey = make_shared<lambertian>(color(0.1, 0.1, 0.1));
    auto light = make_shared<diffuse_light>(color(20, 20, 20));
    world.add(make_shared<sphere>(point3(0, -1005, 0), 1000, grey));
    world.add(make_shared<sphere>(point3(0, 15, 0), 5, light));
    world.add(make_shared<sphere>(point3(-5, 0, 0), 5, make_shared<specular>(color(1.0, 0.1, 0.1), 5)));
    set_frame(cam, cfg, 512, 288, 100);
    cam.max_depth = 10; cfg.depth = 10;
    cam.background = color(0, 0, 0);
    cam.vfov = 90;
    cam.lookfrom = point3(0, 5, -10);
    cam.lookat = point3(0, 0, 0);
    cam.vup = vec3(0, 1, 0);
    cam.focus_dist = (cam.lookat - cam.lookfrom).length() - 2.5;
    cam.defocus_angle = 0;
}

// The reference's OWN inputs (BASELINE configs[3], configs[4]): monkey.obj through mesh::loadObj (32 triangular and
// 468 quad faces -> 968 triangles; both halves of a quad take the UVs of its first three corners, mesh.h:78-81) under
// the transform of main.cpp:399-405 (translate * rotate_y(90) * scale 1.5), textured with Images/earthmap.jpg decoded by
// stb_image and linearised by rtw_image (rtw_stb_image.h:99-121), an earth globe with the same texture (main.cpp:166-167)
// and one point light.  The two files are copied from the reference tree into <asset_dir>/reference/ by the build where
// that tree is mounted (build.ensure_reference_assets); nothing else needs them.
inline void monkey_scene(hittable_list& world, camera& cam, std::vector<point_light>& lights, scene_config& cfg) {
    world.add(make_shared<sphere>(point3(0, -1000, 0), 1000, make_shared<lambertian>(color(0.4, 0.4, 0.4))));
    auto earth_texture = make_shared<image_texture>((cfg.asset_dir + "/reference/earthmap.jpg").c_str());
    auto earth_mat = make_shared<lambertian>(earth_texture);
    {
        mesh monkey;
        glm::mat4 tf = glm::mat4(1.0f);
        tf = glm::translate(tf, glm::vec3(3, 1.5f, 0));
        tf = glm::rotate(tf, glm::radians(90.0f), glm::vec3(0, 1, 0));
        tf = glm::scale(tf, glm::vec3(1.5f, 1.5f, 1.5f));
        monkey.loadObj(cfg.asset_dir + "/reference/monkey.obj", world, earth_mat, tf);
    }
    world.add(make_shared<sphere>(point3(3, 1.2, 3.6), 1.2, earth_mat));
    world.add(make_shared<sphere>(point3(3, 0.8, -3.2), 0.8, make_shared<metal>(color(0.8, 0.8, 0.9), 0.05)));
    lights.push_back(point_light(point3(-6, 8, 2), color(30, 30, 28), 0.5));
    set_frame(cam, cfg, 1920, 1080, 64);
    cam.background = color(0.55, 0.65, 0.85);
    cam.vfov = 30;
    cam.lookfrom = point3(-9, 4, 1.5);
    cam.lookat = point3(3, 1.3, 0.2);
    cam.vup = vec3(0, 1, 0);
    cam.defocus_angle = 0;
}

// main.cpp:128-175 (scene 0): checker ground, dielectric, Perlin sphere, a UV-checkered
// triangle, an image-textured globe, slight defocus
inline void mixed(hittable_list& world, camera& cam, scene_config& cfg) {
    auto checker = make_shared<checker_texture>(0.32, color(0, 0, 0), color(.9, .9, .9));
    world.add(make_shared<sphere>(point3(0, -1000, 0), 1000, make_shared<lambertian>(checker)));
    world.add(make_shared<sphere>(point3(2, 1, 5), 1.0, make_shared<dielectric>(1.5)));
    auto pertext = make_shared<noise_texture>(10);
    world.add(make_shared<sphere>(point3(-2, 1, 5), 1.0, make_shared<lambertian>(pertext)));
    auto checkerT = make_shared<checker_texture_triangle>(0.5, color(0, 0, 0), color(.9, .9, .9));
    world.add(make_shared<triangle>(point3(4, 0, 8), point3(-4, 0, 8), point3(0, 6, 8), make_shared<lambertian>(checkerT)));
    auto earth_texture = make_shared<image_texture>((cfg.asset_dir + "/earth.ppm").c_str());
    world.add(make_shared<sphere>(point3(0, 1, 5), 1.0, make_shared<lambertian>(earth_texture)));
    set_frame(cam, cfg, 512, 288, 100);
    cam.background = color(0.7, 0.8, 1.0);
    cam.vfov = 20;
    cam.lookfrom = point3(1, 4, -10);
    cam.lookat = point3(0, 1, 5);
    cam.vup = vec3(0, 1, 0);
    cam.defocus_angle = 0.1;
    cam.focus_dist = (cam.lookat - cam.lookfrom).length();
}

// Not in main.cpp: a compact scene that puts every primitive, material and texture of
// the hot path in one frame (used by the parity tests so that a single converged render
// exercises all of SURVEY §8a): moving sphere, metal with and without fuzz, dielectric,
// nested checker, marble, image texture, triangles via triangle_quad, a rotated box, a
// smoke sphere (1-object BVH leaf => doubled density), an area light and a
// point light.
inline void kitchen_sink(hittable_list& world, camera& cam, std::vector<point_light>& lights, scene_config& cfg) {
    auto inner = make_shared<checker_texture>(0.5, color(0.8, 0.1, 0.1), color(0.1, 0.1, 0.8));
    auto outer = make_shared<checker_texture>(2.0, inner, make_shared<solid_color>(0.85, 0.85, 0.85));
    // the floor sits at y = -0.013, not 0: a 3-D checker evaluated ON one of its own cell
    // boundaries (floor(inv_scale * 0 +- rounding)) is decided by rounding noise, in the reference
    // as much as here, and would make the scene useless as a parity fixture
    world.add(make_shared<quad>(point3(-20, -0.013, -20), vec3(40, 0, 0), vec3(0, 0, 40), make_shared<lambertian>(outer)));
    world.add(make_shared<quad>(point3(-3, 7, -3), vec3(6, 0, 0), vec3(0, 0, 6), make_shared<diffuse_light>(color(6, 6, 5))));
    world.add(make_shared<sphere>(point3(-4, 1, 0), point3(-4, 1.6, 0), 1.0, make_shared<lambertian>(color(0.7, 0.3, 0.1))));
    world.add(make_shared<sphere>(point3(-1.5, 1, 1.5), 1.0, make_shared<dielectric>(1.5)));
    world.add(make_shared<sphere>(point3(1.2, 1, 2.5), 1.0, make_shared<metal>(color(0.8, 0.8, 0.9), 0.0)));
    world.add(make_shared<sphere>(point3(3.6, 0.8, 0.5), 0.8, make_shared<metal>(color(0.9, 0.6, 0.2), 0.4)));
    world.add(make_shared<sphere>(point3(0.5, 0.7, -1.5), 0.7, make_shared<lambertian>(make_shared<noise_texture>(4))));
    auto earth = make_shared<image_texture>((cfg.asset_dir + "/earth.ppm").c_str());
    world.add(make_shared<sphere>(point3(-2.6, 0.8, -2.2), 0.8, make_shared<lambertian>(earth)));
    auto checkerT = make_shared<checker_texture_triangle>(0.5, color(0.1, 0.3, 0.1), color(.9, .9, .5));
    world.add(triangle_quad(point3(2.0, 0.01, 4.5), 3.0, 3.0, make_shared<lambertian>(checkerT)));
    shared_ptr<hittable> crate = box(point3(0, 0, 0), point3(1.4, 2.2, 1.4), make_shared<lambertian>(color(.73, .73, .73)));
    crate = make_shared<rotate_y>(crate, 25);
    crate = make_shared<translate>(crate, vec3(4.2, 0, 3.0));
    world.add(crate);
    auto smoke_boundary = make_shared<sphere>(point3(-5.5, 1.3, 3.5), 1.3, make_shared<dielectric>(1.5));
    world.add(make_shared<constant_medium>(smoke_boundary, 0.9, color(0.9, 0.9, 0.9)));
    world.add(make_shared<sphere>(point3(6.0, 1.0, -2.0), 1.0, make_shared<specular>(color(0.2, 0.8, 0.3), 3)));
    lights.push_back(point_light(point3(8, 6, -6), color(12, 12, 14), 0.3));

    set_frame(cam, cfg, 480, 270, 256);
    cam.background = color(0.05, 0.06, 0.09);
    cam.vfov = 38;
    cam.lookfrom = point3(1.5, 4.2, -13);
    cam.lookat = point3(0, 1.2, 0.5);
    cam.vup = vec3(0, 1, 0);
    cam.defocus_angle = 0.3;
    cam.focus_dist = 13.5;
}

}  // namespace scenes_detail

inline const char* const* scene_names(int* n) {
    static const char* const names[] = {"book1", "cornell", "cornell_smoke", "mesh", "final",
                                        "quads", "emissive", "specular", "mixed", "kitchen_sink"};
    if (n) *n = (int)(sizeof(names) / sizeof(names[0]));
    return names;
}

// Builds `name` into `world` (seeded), fills the camera and cfg.  Returns false for an
// unknown name.  As in main.cpp:442 the finished world is wrapped in one bvh_node.
inline bool build_scene(const std::string& name, unsigned seed, hittable_list& world, camera& cam,
                        std::vector<point_light>& lights, scene_config& cfg) {
    rtb200::host_srand(seed);
    using namespace scenes_detail;
    if (name == "book1") book1(world, cam, cfg);
    else if (name == "cornell") cornell(world, cam, cfg);
    else if (name == "cornell_smoke") cornell_smoke(world, cam, cfg);
    else if (name == "mesh") mesh_scene(world, cam, lights, cfg);
    else if (name == "final") final_scene(world, cam, cfg);
    else if (name == "quads") quads(world, cam, cfg);
    else if (name == "emissive") emissive(world, cam, cfg);
    else if (name == "specular") specular_scene(world, cam, cfg);
    else if (name == "mixed") mixed(world, cam, cfg);
    else if (name == "kitchen_sink") kitchen_sink(world, cam, lights, cfg);
    else if (name == "monkey") monkey_scene(world, cam, lights, cfg);
    else return false;
    if (cfg.wrap_in_bvh) world = hittable_list(make_shared<bvh_node>(world));
    return true;
}

#endif
