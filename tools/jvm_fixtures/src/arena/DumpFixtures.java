package arena;

import java.io.BufferedWriter;
import java.io.File;
import java.io.FileInputStream;
import java.io.FileWriter;
import java.io.IOException;
import java.util.ArrayList;
import java.util.Collections;
import java.util.IdentityHashMap;
import java.util.Properties;

import org.jbox2d.collision.Manifold;
import org.jbox2d.common.Vec2;
import org.jbox2d.dynamics.Body;
import org.jbox2d.dynamics.Fixture;
import org.jbox2d.dynamics.World;
import org.jbox2d.dynamics.contacts.Contact;

import experiment.ExperimentManager;
import sensors.STCameraImage;

/**
 * Golden-vector dumper for puckswarm_b200's CPU oracle (oracle/): runs the UNMODIFIED reference -- PuckSwarm's own
 * Arena / Robot / controllers on the real JBox2D 2.1.2.2 -- exactly as arena.RunOffline.step does
 * (src/arena/RunOffline.java:72-95) and writes, tick by tick, what tests/test_jvm_fixtures.py compares bit for bit:
 * every body's sweep.c, sweep.a, velocity (float bits), the world contact list (fixture creation indices, manifold
 * point count and ids, accumulated impulses), and robot 0's camera image on think ticks.
 *
 * It lives in package `arena` because Arena.robots / Arena.pucks / Robot.body / Robot.simSuite are package-private.
 * It cannot be compiled in puckswarm_b200's build image (no JDK, no JBox2D jar); tools/jvm_fixtures/README.md has the
 * recipe for a box that has both. JBox2D member names are those of 2.1.2.2.
 *
 * usage: java -cp <classes>:<jars> arena.DumpFixtures <experiment dir> <ticks> <out dir>
 *   <experiment dir> holds common.properties (repetitions = number of seeds) and one or more <code>.properties.
 */
public class DumpFixtures {
    public static void main(String[] args) throws IOException {
        String dir = args[0];
        int ticks = Integer.parseInt(args[1]);
        String outDir = args[2];
        new File(outDir).mkdirs();
        ExperimentManager.activate(false, dir);
        while (true) {
            dumpCurrent(dir, ticks, outDir);
            if (!ExperimentManager.hasNext()) break;
            ExperimentManager.next();
        }
    }

    static void dumpCurrent(String dir, int ticks, String outDir) throws IOException {
        String code = ExperimentManager.getCurrent().getStringCodeWithoutSeed();
        int seed = ExperimentManager.getCurrent().getIndex();
        // RunOffline.recreateArena (:97-113)
        Robot.nameIndex = 0;
        World world = new World(new Vec2(0, 0), false);
        Arena arena = new Arena(world, null, false);
        IdentityHashMap<Fixture, Integer> proxyOf = fixtureCreationIndices(world);
        ArrayList<Body> bodies = new ArrayList<Body>();
        for (Puck p : arena.pucks) bodies.add(p.body);
        for (Robot r : arena.robots) bodies.add(r.body);

        Properties props = new Properties();
        props.load(new FileInputStream(dir + File.separatorChar + "common.properties"));
        Properties own = new Properties();
        own.load(new FileInputStream(dir + File.separatorChar + code + ".properties"));
        props.putAll(own);

        BufferedWriter out = new BufferedWriter(new FileWriter(outDir + File.separatorChar + "jvm_" + code + "_" + seed + ".json"));
        out.write("{\"format\": 1, \"source\": \"BOTSlab/PuckSwarm on JBox2D 2.1.2.2 (tools/jvm_fixtures/src/arena/DumpFixtures.java)\",\n");
        out.write(" \"code\": \"" + code + "\", \"seed\": " + seed + ", \"n_pucks\": " + arena.pucks.size() + ", \"n_robots\": " + arena.robots.size() + ",\n");
        out.write(" \"properties\": {");
        boolean first = true;
        for (String k : props.stringPropertyNames()) {
            out.write((first ? "" : ", ") + "\"" + k + "\": \"" + props.getProperty(k) + "\"");
            first = false;
        }
        out.write("},\n \"states\": [\n");
        writeState(out, 0, world, bodies, proxyOf, null);
        int updateCount = 0;
        for (int t = 0; t < ticks; t++) {
            STCameraImage image = null;
            // RunOffline.step (:87-94)
            if (updateCount % PuckSwarmTest.STEP_INTERVAL == 0) {
                arena.step(true, true, true, false, 0, false);
                image = arena.robots.get(0).simSuite.getCameraImage();
            } else {
                arena.coast(true, true, true);
            }
            updateCount++;
            world.step(1 / 14f, 3, 100);
            out.write(",\n");
            writeState(out, updateCount, world, bodies, proxyOf, image);
        }
        out.write("\n]}\n");
        out.close();
        System.out.println("wrote jvm_" + code + "_" + seed + ".json (" + ticks + " ticks)");
    }

    /** Fixture -> creation index over the whole world = broad-phase proxy id order (bodies and fixtures are linked newest first). */
    static IdentityHashMap<Fixture, Integer> fixtureCreationIndices(World world) {
        ArrayList<Body> bodies = new ArrayList<Body>();
        for (Body b = world.getBodyList(); b != null; b = b.getNext()) bodies.add(b);
        Collections.reverse(bodies);
        IdentityHashMap<Fixture, Integer> map = new IdentityHashMap<Fixture, Integer>();
        int next = 0;
        for (Body b : bodies) {
            ArrayList<Fixture> fs = new ArrayList<Fixture>();
            for (Fixture f = b.getFixtureList(); f != null; f = f.getNext()) fs.add(f);
            Collections.reverse(fs);
            for (Fixture f : fs) map.put(f, next++);
        }
        return map;
    }

    static void writeState(BufferedWriter out, int tick, World world, ArrayList<Body> bodies, IdentityHashMap<Fixture, Integer> proxyOf,
                           STCameraImage image) throws IOException {
        StringBuilder sb = new StringBuilder();
        sb.append("  {\"tick\": ").append(tick).append(", \"body\": [");
        for (int i = 0; i < bodies.size(); i++) {
            Body b = bodies.get(i);
            Vec2 v = b.getLinearVelocity();
            if (i > 0) sb.append(", ");
            sb.append(bits(b.m_sweep.c.x)).append(", ").append(bits(b.m_sweep.c.y)).append(", ").append(bits(b.m_sweep.a)).append(", ")
              .append(bits(v.x)).append(", ").append(bits(v.y)).append(", ").append(bits(b.getAngularVelocity()));
        }
        sb.append("],\n   \"contacts\": [");
        // the world contact list is newest first; the exchange format keeps creation order
        ArrayList<Contact> cs = new ArrayList<Contact>();
        for (Contact c = world.getContactList(); c != null; c = c.getNext()) cs.add(c);
        Collections.reverse(cs);
        for (int i = 0; i < cs.size(); i++) {
            Contact c = cs.get(i);
            Manifold m = c.getManifold();
            if (i > 0) sb.append(", ");
            sb.append("[").append(proxyOf.get(c.getFixtureA())).append(", ").append(proxyOf.get(c.getFixtureB())).append(", ").append(m.pointCount);
            for (int j = 0; j < 2; j++) {
                boolean live = j < m.pointCount;
                int id = 0;
                if (live) {
                    org.jbox2d.collision.ContactID.Features f = m.points[j].id.features;
                    id = (f.referenceEdge & 7) | ((f.incidentEdge & 7) << 3) | ((f.incidentVertex & 1) << 6) | ((f.flip & 1) << 7);   // ContactID::pack of oracle/jb2d.h
                }
                sb.append(", ").append(id).append(", ").append(live ? bits(m.points[j].normalImpulse) : 0).append(", ").append(live ? bits(m.points[j].tangentImpulse) : 0);
            }
            sb.append("]");
        }
        sb.append("]");
        if (image != null) {
            // SensedType ordinals, column-major like STImage.pixels[x][y], run-length encoded as [value, run, value, run, ...]
            sb.append(",\n   \"camera_robot0_rle\": [");
            int prev = -1, run = 0;
            boolean firstRun = true;
            for (int x = 0; x < image.width; x++)
                for (int y = 0; y < image.height; y++) {
                    int v = image.pixels[x][y].ordinal();
                    if (v == prev) { run++; continue; }
                    if (prev >= 0) { sb.append(firstRun ? "" : ", ").append(prev).append(", ").append(run); firstRun = false; }
                    prev = v; run = 1;
                }
            sb.append(firstRun ? "" : ", ").append(prev).append(", ").append(run).append("]");
        }
        sb.append("}");
        out.write(sb.toString());
    }

    static int bits(float f) { return Float.floatToRawIntBits(f); }
}
