package cc.mrlda.b200;

import java.io.IOException;
import java.util.ArrayList;
import java.util.Iterator;
import java.util.List;
import java.util.Set;

import org.apache.hadoop.conf.Configuration;
import org.apache.hadoop.fs.FileSystem;
import org.apache.hadoop.fs.Path;
import org.apache.hadoop.io.IOUtils;
import org.apache.hadoop.io.SequenceFile;

import cc.mrlda.InformedPrior;
import edu.umd.cloud9.util.map.HMapIV;

/**
 * -informedprior: the eta file InformedPrior writes (SequenceFile&lt;IntWritable,
 * ArrayListOfIntsWritable&gt;, topic -&gt; seed terms, InformedPrior.java:126-177) flattened into the
 * (topic, term) pairs mrlda_set_informed_prior takes. Replaces TermReducer.configure's importEta call
 * (TermReducer.java:95-105).
 */
public final class EtaFile {
  private EtaFile() {
  }

  public static void load(FileSystem fs, Path etaFile, Configuration conf, MrLdaB200 gpu)
      throws IOException {
    HMapIV<Set<Integer>> lambdaMap;
    SequenceFile.Reader reader = null;
    try {
      reader = new SequenceFile.Reader(fs, etaFile, conf);
      lambdaMap = InformedPrior.importEta(reader);
    } finally {
      IOUtils.closeStream(reader);
    }
    List<Integer> topics = new ArrayList<Integer>();
    List<Integer> terms = new ArrayList<Integer>();
    Iterator<Integer> topicItr = lambdaMap.keySet().iterator();
    while (topicItr.hasNext()) {
      int topic = topicItr.next();
      for (int term : lambdaMap.get(topic)) {
        topics.add(topic);
        terms.add(term);
      }
    }
    int[] topic = new int[topics.size()];
    int[] term = new int[terms.size()];
    for (int i = 0; i < topic.length; i++) {
      topic[i] = topics.get(i);
      term[i] = terms.get(i);
    }
    gpu.setInformedPrior(topic, term);
  }
}
