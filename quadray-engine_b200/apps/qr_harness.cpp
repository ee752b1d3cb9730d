/*
 * qr_harness.cpp: headless caller of the rt_Platform / rt_Scene public API.
 *
 * It plays the role RooT (root/RooT.h:1014-1121, 590-631) and core_test
 * (test/core_test.cpp:939-1046) play for the reference: construct a platform
 * with a pthread pool, construct a scene from the statically-linkable scene
 * data, call scene->render(time) N times, read get_frame().  It only uses the
 * public API (core/engine/engine.h:131-152, 325-359), so the very same source
 * links against
 *   - the unmodified reference core (oracle/_ref/qr_ref_harness: parity
 *     goldens and the CPU baseline), and
 *   - the reference engine + this repo's replacement tracer TU
 *     (build/qr_b200_harness: the drop-in B200 path).
 *
 * RooT itself needs an X display even with -o (root/RooT_linux.cpp:194-199),
 * which is why a headless caller exists at all.
 */

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <pthread.h>
#include <sched.h>
#include <sys/time.h>
#include <sys/mman.h>
#include <vector>
#include <algorithm>

#include "engine.h"

#include "all_scn.h"
#include "scn_test01.h"
#include "scn_test02.h"
#include "scn_test03.h"
#include "scn_test04.h"
#include "scn_test05.h"
#include "scn_test06.h"
#include "scn_test07.h"
#include "scn_test08.h"
#include "scn_test09.h"
#include "scn_test10.h"
#include "scn_test11.h"
#include "scn_test12.h"
#include "scn_test13.h"
#include "scn_test14.h"
#include "scn_test15.h"
#include "scn_test16.h"
#include "scn_test17.h"
#include "scn_test18.h"

#include "qr_synth_scene.h"

struct SceneEntry { const char *name; rt_SCENE *root; };

static SceneEntry g_scenes[] =
{
    { "test01", &scn_test01::sc_root }, { "test02", &scn_test02::sc_root },
    { "test03", &scn_test03::sc_root }, { "test04", &scn_test04::sc_root },
    { "test05", &scn_test05::sc_root }, { "test06", &scn_test06::sc_root },
    { "test07", &scn_test07::sc_root }, { "test08", &scn_test08::sc_root },
    { "test09", &scn_test09::sc_root }, { "test10", &scn_test10::sc_root },
    { "test11", &scn_test11::sc_root }, { "test12", &scn_test12::sc_root },
    { "test13", &scn_test13::sc_root }, { "test14", &scn_test14::sc_root },
    { "test15", &scn_test15::sc_root }, { "test16", &scn_test16::sc_root },
    { "test17", &scn_test17::sc_root }, { "test18", &scn_test18::sc_root },
    { "demo01", &scn_demo01::sc_root }, { "demo02", &scn_demo02::sc_root },
    { "demo03", &scn_demo03::sc_root },
};

/* ---------------------------------------------------------------- memory -- */

static pthread_mutex_t g_alloc_mtx = PTHREAD_MUTEX_INITIALIZER;

static rt_pntr sys_alloc(rt_size size)
{
    pthread_mutex_lock(&g_alloc_mtx);
    void *p = mmap(NULL, size, PROT_READ | PROT_WRITE,
                   MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    pthread_mutex_unlock(&g_alloc_mtx);
    if (p == MAP_FAILED || p == NULL)
    {
        throw rt_Exception("alloc failed in qr_harness sys_alloc");
    }
    return p;
}

static rt_void sys_free(rt_pntr ptr, rt_size size)
{
    pthread_mutex_lock(&g_alloc_mtx);
    munmap(ptr, size);
    pthread_mutex_unlock(&g_alloc_mtx);
}

static double now_ms()
{
    timeval tm;
    gettimeofday(&tm, NULL);
    return tm.tv_sec * 1000.0 + tm.tv_usec / 1000.0;
}

/* ----------------------------------------------------------- thread pool -- */
/*
 * Same contract as the reference's platform pool (engine.h:71-74; shape of
 * root/RooT_linux.cpp:546-793): f_init makes "thnum" workers and reports the
 * usable count back through set_thnum, f_update / f_render hand one phase to
 * every worker (worker i runs update_slice(i) / render_slice(i)) and block
 * until all are done.  Implemented with a generation counter + condvars.
 */
struct Pool
{
    rt_Platform        *pfm;
    int                 thnum;
    pthread_t          *thr;
    pthread_mutex_t     mtx;
    pthread_cond_t      go, done;
    unsigned            gen;
    int                 cmd, phase, pending, quit;
    const char         *err;
};

struct Worker { Pool *pool; int index; };

static void *pool_worker(void *arg)
{
    Worker *w = (Worker *)arg;
    Pool *p = w->pool;
    unsigned seen = 0;

    for (;;)
    {
        pthread_mutex_lock(&p->mtx);
        while (p->gen == seen && !p->quit)
        {
            pthread_cond_wait(&p->go, &p->mtx);
        }
        if (p->quit)
        {
            pthread_mutex_unlock(&p->mtx);
            break;
        }
        seen = p->gen;
        int cmd = p->cmd, phase = p->phase;
        pthread_mutex_unlock(&p->mtx);

        try
        {
            rt_Scene *scn = p->pfm->get_cur_scene();
            if (cmd == 1) scn->update_slice(w->index, phase);
            if (cmd == 2) scn->render_slice(w->index, phase);
        }
        catch (rt_Exception e)
        {
            p->err = e.err;
        }

        pthread_mutex_lock(&p->mtx);
        if (--p->pending == 0)
        {
            pthread_cond_signal(&p->done);
        }
        pthread_mutex_unlock(&p->mtx);
    }
    delete w;
    return NULL;
}

static rt_pntr pool_init(rt_si32 thnum, rt_Platform *pfm)
{
    bool feedback = thnum > 0;
    thnum = thnum < 0 ? -thnum : thnum;

    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    sched_getaffinity(0, sizeof(allowed), &allowed);
    int ncpu = CPU_COUNT(&allowed);
    if (feedback && thnum > ncpu)
    {
        thnum = ncpu;
    }

    Pool *p = new Pool();
    p->pfm = pfm;
    p->thnum = thnum;
    p->thr = new pthread_t[thnum];
    pthread_mutex_init(&p->mtx, NULL);
    pthread_cond_init(&p->go, NULL);
    pthread_cond_init(&p->done, NULL);
    p->gen = 0; p->cmd = 0; p->phase = 0; p->pending = 0; p->quit = 0;
    p->err = NULL;

    int cpu = -1;
    for (int i = 0; i < thnum; i++)
    {
        Worker *w = new Worker();
        w->pool = p; w->index = i;
        pthread_create(&p->thr[i], NULL, pool_worker, w);
        /* one worker per allowed cpu, round-robin */
        do { cpu = (cpu + 1) % CPU_SETSIZE; } while (!CPU_ISSET(cpu, &allowed));
        cpu_set_t one;
        CPU_ZERO(&one);
        CPU_SET(cpu, &one);
        pthread_setaffinity_np(p->thr[i], sizeof(one), &one);
    }
    if (feedback)
    {
        pfm->set_thnum(thnum);
    }
    return p;
}

static rt_void pool_term(rt_pntr tdata, rt_si32 thnum)
{
    Pool *p = (Pool *)tdata;
    pthread_mutex_lock(&p->mtx);
    p->quit = 1;
    pthread_cond_broadcast(&p->go);
    pthread_mutex_unlock(&p->mtx);
    for (int i = 0; i < p->thnum; i++)
    {
        pthread_join(p->thr[i], NULL);
    }
    delete[] p->thr;
    delete p;
}

static void pool_run(Pool *p, int cmd, int phase)
{
    pthread_mutex_lock(&p->mtx);
    p->cmd = cmd; p->phase = phase; p->pending = p->thnum;
    p->gen++;
    pthread_cond_broadcast(&p->go);
    while (p->pending != 0)
    {
        pthread_cond_wait(&p->done, &p->mtx);
    }
    pthread_mutex_unlock(&p->mtx);
    if (p->err != NULL)
    {
        const char *e = p->err;
        p->err = NULL;
        throw rt_Exception(e);
    }
}

static rt_void pool_update(rt_pntr tdata, rt_si32 thnum, rt_si32 phase)
{
    pool_run((Pool *)tdata, 1, phase);
}

static rt_void pool_render(rt_pntr tdata, rt_si32 thnum, rt_si32 phase)
{
    pool_run((Pool *)tdata, 2, phase);
}

/* ------------------------------------------------------------------ main -- */

static void usage()
{
    printf("qr_harness -s <test01..test18|demo01..demo03|synth> [-x w] [-y h]\n"
           "  [synth: -N quadrics -S seed -E extent -R 0|1 (rotations) -M metal_permille]\n"
           "  [-a 0|1|2 (fsaa none/2x/4x)] [-p none|full|default|0xHEX (opts)]\n"
           "  [-g (gamma prop on)] [-r (fresnel prop on)] [-c cam_idx]\n"
           "  [-t threads (0 = stub, sequential)] [-f frames] [-w warmup]\n"
           "  [-b time_begin_ms] [-d time_delta_ms] [-n simd -k size -v type]\n"
           "  [-u (freeze update after 1st frame: RT_OPTS_UPDATE_EXT0)]\n"
           "  [-o out.raw (last frame, x_res*y_res u32)] [-q (quiet)]\n"
           "  [-T t.raw (oracle/_ref/qr_ref_tdump only: primary hit distance per sample)]\n"
           "  [-Q (path tracer, rt_Scene::set_pton: -f frames accumulate, use -d 0)]\n");
}

int main(int argc, char **argv)
{
    const char *scene_name = "test01", *out = NULL, *opts_s = "default", *t_out = NULL;
    int x_res = 800, y_res = 480, fsaa = 0, threads = 0, frames = 1, warm = 0;
    int cam_idx = 0, n_simd = 0, k_size = 0, s_type = 0, quiet = 0, pt_mode = 0;
    int gamma_on = 0, fresnel_on = 0, freeze = 0;
    long t_begin = 0, t_delta = 16;
    qr_synth::Params synth = { 1000, 1, 100.0f, 1, 0 };

    for (int i = 1; i < argc; i++)
    {
        const char *a = argv[i];
        const char *v = i + 1 < argc ? argv[i + 1] : "";
        if      (!strcmp(a, "-s")) { scene_name = v; i++; }
        else if (!strcmp(a, "-x")) { x_res = atoi(v); i++; }
        else if (!strcmp(a, "-y")) { y_res = atoi(v); i++; }
        else if (!strcmp(a, "-a")) { fsaa = atoi(v); i++; }
        else if (!strcmp(a, "-p")) { opts_s = v; i++; }
        else if (!strcmp(a, "-g")) { gamma_on = 1; }
        else if (!strcmp(a, "-r")) { fresnel_on = 1; }
        else if (!strcmp(a, "-u")) { freeze = 1; }
        else if (!strcmp(a, "-c")) { cam_idx = atoi(v); i++; }
        else if (!strcmp(a, "-t")) { threads = atoi(v); i++; }
        else if (!strcmp(a, "-f")) { frames = atoi(v); i++; }
        else if (!strcmp(a, "-w")) { warm = atoi(v); i++; }
        else if (!strcmp(a, "-b")) { t_begin = atol(v); i++; }
        else if (!strcmp(a, "-d")) { t_delta = atol(v); i++; }
        else if (!strcmp(a, "-n")) { n_simd = atoi(v) / 128; i++; }
        else if (!strcmp(a, "-k")) { k_size = atoi(v); i++; }
        else if (!strcmp(a, "-v")) { s_type = atoi(v); i++; }
        else if (!strcmp(a, "-o")) { out = v; i++; }
        else if (!strcmp(a, "-q")) { quiet = 1; }
        else if (!strcmp(a, "-T")) { t_out = v; i++; }
        else if (!strcmp(a, "-Q")) { pt_mode = 1; }
        else if (!strcmp(a, "-N")) { synth.n = atoi(v); i++; }
        else if (!strcmp(a, "-S")) { synth.seed = (unsigned)atol(v); i++; }
        else if (!strcmp(a, "-E")) { synth.extent = (float)atof(v); i++; }
        else if (!strcmp(a, "-R")) { synth.rotate = atoi(v); i++; }
        else if (!strcmp(a, "-M")) { synth.metal = atoi(v); i++; }
        else { usage(); return 2; }
    }

    rt_SCENE root;
    bool found = false;
    for (size_t i = 0; i < sizeof(g_scenes) / sizeof(g_scenes[0]); i++)
    {
        if (!strcmp(g_scenes[i].name, scene_name))
        {
            root = *g_scenes[i].root;
            found = true;
        }
    }
    if (!found && !strcmp(scene_name, "synth") && synth.n > 0)
    {
        root = qr_synth::build(synth);
        found = true;
    }
    if (!found)
    {
        fprintf(stderr, "unknown scene %s\n", scene_name);
        return 2;
    }
    /* scene data lists optimisations to turn OFF; turning RT_OPTS_GAMMA /
     * RT_OPTS_FRESNEL off enables the props (format.h:59-60, 73-75) */
    if (gamma_on)   root.opts |= RT_OPTS_GAMMA;
    if (fresnel_on) root.opts |= RT_OPTS_FRESNEL;

    int rc = 0;
    try
    {
        rt_Platform *pfm = threads > 0 ?
            new rt_Platform(sys_alloc, sys_free, threads,
                            pool_init, pool_term, pool_update, pool_render) :
            new rt_Platform(sys_alloc, sys_free);

        int simd = pfm->set_simd(simd_init(n_simd, s_type, k_size));
        fsaa = pfm->set_fsaa(fsaa);
        int tile_w = pfm->get_tile_w();

        int x_row = (x_res + RT_SIMD_WIDTH - 1) & ~(RT_SIMD_WIDTH - 1);

        rt_Scene *scene = new(pfm) rt_Scene(&root, x_res, y_res, x_row,
                                            RT_NULL, pfm);
        pfm->set_cur_scene(scene);

        for (int c = cam_idx; c > 0; c--)
        {
            scene->next_cam();
        }

        if (!strcmp(opts_s, "none"))
        {
            scene->set_opts(RT_OPTS_NONE);
        }
        else if (!strcmp(opts_s, "full"))
        {
            scene->set_opts(RT_OPTS_FULL);
        }
        else if (strcmp(opts_s, "default"))
        {
            scene->set_opts((rt_si32)strtoul(opts_s, NULL, 0));
        }

        if (pt_mode && !scene->set_pton(1))
        {
            fprintf(stderr, "the scene does not allow the path tracer (RT_OPTS_PT)\n");
            return 2;
        }

        std::vector<double> ms;
        for (int j = 0; j < warm + frames; j++)
        {
            double t0 = now_ms();
            scene->render((rt_time)(t_begin + t_delta * j));
            double t1 = now_ms();
            if (j >= warm)
            {
                ms.push_back(t1 - t0);
            }
            if (freeze && j == 0)
            {
                scene->set_opts(scene->get_opts() | RT_OPTS_UPDATE_EXT0);
            }
        }

        if (out != NULL)
        {
            FILE *f = fopen(out, "wb");
            if (f == NULL)
            {
                fprintf(stderr, "cannot open %s\n", out);
                rc = 1;
            }
            else
            {
                rt_ui32 *fr = scene->get_frame();
                int row = scene->get_x_row();
                for (int y = 0; y < y_res; y++)
                {
                    fwrite(fr + (size_t)y * row, 4, x_res, f);
                }
                fclose(f);
            }
        }

        if (t_out != NULL)
        {
#ifdef QR_TDUMP
            /* the patched reference of oracle/Makefile's tdump target left
             * ctx_T_BUF(0) of every packet in the red plane:
             * (y * x_row + x) << fsaa floats in, lanes in packet order */
            FILE *f = fopen(t_out, "wb");
            rt_real *tb = scene->qr_tbuf();
            if (f == NULL || tb == NULL)
            {
                fprintf(stderr, "cannot dump T to %s\n", t_out);
                rc = 1;
            }
            else
            {
                int row = scene->get_x_row();
                for (int y = 0; y < y_res; y++)
                {
                    fwrite(tb + (((size_t)y * row) << fsaa), sizeof(rt_real), (size_t)x_res << fsaa, f);
                }
            }
            if (f != NULL) fclose(f);
#else
            fprintf(stderr, "-T needs the tdump build (oracle/Makefile: make tdump)\n");
            rc = 2;
#endif
        }

        std::vector<double> s = ms;
        std::sort(s.begin(), s.end());
        double sum = 0;
        for (size_t i = 0; i < s.size(); i++) sum += s[i];
        if (!quiet || true)
        {
            printf("{\"scene\": \"%s\", \"x_res\": %d, \"y_res\": %d, "
                   "\"fsaa\": %d, \"opts\": \"0x%08X\", \"threads\": %d, "
                   "\"simd\": \"%dx%dv%d\", \"tile_w\": %d, \"frames\": %d, "
                   "\"ms_min\": %.3f, \"ms_med\": %.3f, \"ms_mean\": %.3f, "
                   "\"ms_total\": %.3f}\n",
                   scene_name, x_res, y_res, fsaa, (unsigned)scene->get_opts(),
                   pfm->get_thnum(),
                   (simd & 0xFF) * 128, (simd >> 16) & 0xFF, (simd >> 8) & 0xFF,
                   tile_w, (int)s.size(),
                   s.empty() ? 0.0 : s[0],
                   s.empty() ? 0.0 : s[s.size() / 2],
                   s.empty() ? 0.0 : sum / s.size(), sum);
        }

        delete scene;
        delete pfm;
    }
    catch (rt_Exception e)
    {
        fprintf(stderr, "Exception: %s\n", e.err);
        rc = 1;
    }
    return rc;
}
