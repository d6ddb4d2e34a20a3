/*
 * RandomizerState.java -- save / restore of BEAST's global Randomizer so that a window of
 * uniforms can be "peeked": draw the window, run the kernel, restore the saved state and advance
 * it by exactly the number of doubles the kernel consumed.  BEAST's Randomizer wraps a
 * MersenneTwisterFast whose state is reachable by reflection (field "random") [EXT]; the restated
 * generator in csrc/host/randomizer.cpp (mtt_rng_peek_doubles / mtt_rng_skip_doubles) is the
 * reference for the intended semantics.  NOT compiled here (no JDK).
 */
package b200mtt;

import beast.base.util.MersenneTwisterFast;
import beast.base.util.Randomizer;

import java.lang.reflect.Field;

final class RandomizerState {
    private static final Field RANDOM;

    static {
        try {
            RANDOM = Randomizer.class.getDeclaredField("random");
            RANDOM.setAccessible(true);
        } catch (ReflectiveOperationException e) {
            throw new ExceptionInInitializerError(e);
        }
    }

    private static MersenneTwisterFast snapshot;

    static long[] save() {
        try {
            snapshot = (MersenneTwisterFast) ((MersenneTwisterFast) RANDOM.get(null)).clone();
            return null;
        } catch (ReflectiveOperationException | CloneNotSupportedException e) {
            throw new IllegalStateException(e);
        }
    }

    static void restoreAndSkip(long[] ignored, long consumed) {
        try {
            RANDOM.set(null, snapshot);
            for (long i = 0; i < consumed; i++)
                Randomizer.nextDouble();
        } catch (ReflectiveOperationException e) {
            throw new IllegalStateException(e);
        }
    }
}
