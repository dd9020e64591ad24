package cc.mrlda.b200;

import java.io.IOException;
import java.util.Iterator;

import org.apache.hadoop.conf.Configuration;
import org.apache.hadoop.fs.FileSystem;
import org.apache.hadoop.fs.Path;
import org.apache.hadoop.io.IOUtils;
import org.apache.hadoop.io.SequenceFile;

import com.google.common.base.Preconditions;

import edu.umd.cloud9.io.map.HMapIDW;
import edu.umd.cloud9.io.pair.PairOfIntFloat;

/**
 * beta-N files: SequenceFile&lt;PairOfIntFloat, HMapIDW&gt;, one record per topic, key = (topic 1..K,
 * (float) digamma(sum lambda)), value = term -&gt; digamma(lambda) (TermReducer.java:173-195,232-235;
 * read by DocumentMapper.importBeta, DocumentMapper.java:475-536).
 */
public final class BetaFile {
  private BetaFile() {
  }

  /** Loads a beta file into the flat arrays mrlda_set_beta takes and hands them to the GPU. */
  public static void load(FileSystem fs, Path betaFile, Configuration conf, MrLdaB200 gpu)
      throws IOException {
    final int K = gpu.getNumberOfTopics(), V = gpu.getNumberOfTerms();
    double[] digammaLambda = new double[V * K];
    float[] normaliser = new float[K];
    byte[] present = new byte[V];
    PairOfIntFloat key = new PairOfIntFloat();
    HMapIDW value = new HMapIDW();
    SequenceFile.Reader reader = null;
    try {
      reader = new SequenceFile.Reader(fs, betaFile, conf);
      while (reader.next(key, value)) {
        Preconditions.checkArgument(key.getLeftElement() > 0 && key.getLeftElement() <= K,
            "Invalid beta vector for term " + key.getLeftElement() + "...");
        int k = key.getLeftElement() - 1;
        normaliser[k] = key.getRightElement();
        Iterator<Integer> itr = value.keySet().iterator();
        while (itr.hasNext()) {
          int term = itr.next();
          Preconditions.checkArgument(term > 0 && term <= V, "Invalid term index " + term + "...");
          digammaLambda[(term - 1) * K + k] = value.get(term);
          present[term - 1] = 1;
        }
      }
    } finally {
      IOUtils.closeStream(reader);
    }
    gpu.setBeta(digammaLambda, normaliser, present);
  }

  /** Writes the GPU's current beta as the file the next iteration (or -test) imports. */
  public static void write(FileSystem fs, Path betaFile, Configuration conf, MrLdaB200 gpu)
      throws IOException {
    final int K = gpu.getNumberOfTopics(), V = gpu.getNumberOfTerms();
    double[] digammaLambda = new double[V * K];
    float[] normaliser = new float[K];
    byte[] present = new byte[V];
    gpu.getBeta(digammaLambda, normaliser, present);
    PairOfIntFloat key = new PairOfIntFloat();
    HMapIDW value = new HMapIDW();
    SequenceFile.Writer writer = null;
    try {
      writer = new SequenceFile.Writer(fs, conf, betaFile, PairOfIntFloat.class, HMapIDW.class);
      for (int k = 0; k < K; k++) {
        value.clear();
        for (int v = 0; v < V; v++) {
          if (present[v] != 0) {
            value.put(v + 1, digammaLambda[v * K + k]);
          }
        }
        key.set(k + 1, normaliser[k]);
        writer.append(key, value);
      }
    } finally {
      IOUtils.closeStream(writer);
    }
  }
}
