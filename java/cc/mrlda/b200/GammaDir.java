package cc.mrlda.b200;

import java.io.IOException;

import org.apache.hadoop.conf.Configuration;
import org.apache.hadoop.fs.FileSystem;
import org.apache.hadoop.fs.Path;
import org.apache.hadoop.io.IOUtils;
import org.apache.hadoop.io.IntWritable;
import org.apache.hadoop.io.SequenceFile;

import cc.mrlda.Document;
import edu.umd.cloud9.util.map.HMapII;

/**
 * gamma-N directories: the "gamma" named output of the map phase, SequenceFile&lt;IntWritable,
 * Document&gt; with the updated gamma embedded in every document (DocumentMapper.java:342-346), which
 * VariationalInference renames into outputPath/gamma-(N+1) (VariationalInference.java:359-379).
 */
public final class GammaDir {
  private GammaDir() {
  }

  /** Writes dir/gamma_gamma-m-RRRRR for this rank from the CSR corpus and the GPU's gamma. */
  public static void write(FileSystem fs, Path dir, Configuration conf, CsrCorpus csr, MrLdaB200 gpu,
      int rank) throws IOException {
    final int K = gpu.getNumberOfTopics(), D = csr.numberOfDocuments;
    double[] gamma = new double[D * K];
    gpu.getGamma(gamma);
    fs.mkdirs(dir);
    Path part = new Path(dir, String.format("gamma_gamma-m-%05d", rank));
    IntWritable key = new IntWritable();
    SequenceFile.Writer writer = null;
    try {
      writer = new SequenceFile.Writer(fs, conf, part, IntWritable.class, Document.class);
      for (int d = 0; d < D; d++) {
        if (csr.rowPtr[d + 1] == csr.rowPtr[d]) {
          continue; // content == null: the mapper returns before the gamma output (DocumentMapper.java:197-201)
        }
        HMapII content = new HMapII();
        for (long z = csr.rowPtr[d]; z < csr.rowPtr[d + 1]; z++) {
          content.put(csr.termId[(int) z], csr.count[(int) z]);
        }
        Document doc = new Document(content);
        double[] g = new double[K];
        System.arraycopy(gamma, d * K, g, 0, K);
        doc.setGamma(g);
        key.set(csr.docId[d]);
        writer.append(key, doc);
      }
    } finally {
      IOUtils.closeStream(writer);
    }
  }
}
